// ipc_probe.cu -- feasibility probe for the multi-GPU halo path: two PROCESSES (one per GPU) map each other's
// device buffers with CUDA IPC, write into them from a kernel and hand-shake through flags in peer memory.
// Prints the round-trip latency of a kernel-level flag ping-pong and the bandwidth of peer stores.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 scripts/ipc_probe.cu -o gpurun_out/ipc_probe && gpurun_out/ipc_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("rank %d: %s failed: %s\n", rank, #x, cudaGetErrorString(e)); exit(2); } } while (0)

__global__ void k_signal(volatile int *peer_flag, int v) { __threadfence_system(); *peer_flag = v; }
__global__ void k_wait(volatile int *flag, int v) { while (*flag < v) { } __threadfence_system(); }
__global__ void k_store(int4 *peer, int n, int v) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) peer[i] = make_int4(v, i, v, i);
}
__global__ void k_check(const int4 *buf, int n, int v, int *bad) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (buf[i].x != v || buf[i].y != i) atomicAdd(bad, 1);
}

int main() {
  int p01[2], p10[2];
  if (pipe(p01) || pipe(p10)) return 1;
  pid_t pid = fork();          // before any CUDA call: a CUDA context does not survive fork
  const int rank = pid == 0 ? 1 : 0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 2) { printf("need 2 GPUs, have %d\n", ndev); return 1; }
  const int rfd = rank == 0 ? p10[0] : p01[0], wfd = rank == 0 ? p01[1] : p10[1];
  CK(cudaSetDevice(rank));
  const int n = 1 << 20;   // 16 MB of int4
  int4 *buf; int *flag, *bad;
  CK(cudaMalloc(&buf, sizeof(int4) * n));
  CK(cudaMalloc(&flag, 256));
  CK(cudaMalloc(&bad, 4));
  CK(cudaMemset(flag, 0, 256));
  CK(cudaMemset(bad, 0, 4));
  CK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t mine[2], theirs[2];
  CK(cudaIpcGetMemHandle(&mine[0], buf));
  CK(cudaIpcGetMemHandle(&mine[1], flag));
  if (write(wfd, mine, sizeof mine) != (ssize_t)sizeof mine) return 3;
  if (read(rfd, theirs, sizeof theirs) != (ssize_t)sizeof theirs) return 3;
  int4 *pbuf; int *pflag;
  CK(cudaIpcOpenMemHandle((void **)&pbuf, theirs[0], cudaIpcMemLazyEnablePeerAccess));
  CK(cudaIpcOpenMemHandle((void **)&pflag, theirs[1], cudaIpcMemLazyEnablePeerAccess));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  // 1. flag ping-pong: rank 0 signals 2k+1, rank 1 answers 2k+2
  const int iters = 2000;
  CK(cudaEventRecord(e0, st));
  for (int k = 0; k < iters; k++) {
    if (rank == 0) { k_signal<<<1, 1, 0, st>>>(pflag, 2 * k + 1); k_wait<<<1, 1, 0, st>>>(flag, 2 * k + 2); }
    else { k_wait<<<1, 1, 0, st>>>(flag, 2 * k + 1); k_signal<<<1, 1, 0, st>>>(pflag, 2 * k + 2); }
  }
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("rank %d: flag ping-pong round trip %.2f us (direct launches)\n", rank, 1000.f * ms / iters);
  // 2. peer stores + handshake + check
  CK(cudaEventRecord(e0, st));
  const int reps = 20;
  for (int r = 0; r < reps; r++) k_store<<<592, 256, 0, st>>>(pbuf, n, 7 + rank);
  CK(cudaEventRecord(e1, st));
  k_signal<<<1, 1, 0, st>>>(pflag + 1, 1);
  k_wait<<<1, 1, 0, st>>>(flag + 1, 1);
  k_check<<<592, 256, 0, st>>>(buf, n, 7 + (rank ^ 1), bad);
  CK(cudaStreamSynchronize(st));
  CK(cudaEventElapsedTime(&ms, e0, e1));
  int hbad; CK(cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost));
  printf("rank %d: peer store %.1f GB/s, %d bad entries after handshake\n", rank, reps * 16.0 * n / 1e9 / (ms * 1e-3), hbad);
  // 3. the same ping-pong replayed from a CUDA graph (spin-wait kernels as graph nodes)
  cudaGraph_t g; cudaGraphExec_t gx;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
  // flags continue from 2*iters
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphDestroy(g));
  int can = 0;
  CK(cudaDeviceCanAccessPeer(&can, rank, rank ^ 1));
  printf("rank %d: cudaDeviceCanAccessPeer = %d\n", rank, can);
  CK(cudaIpcCloseMemHandle(pbuf)); CK(cudaIpcCloseMemHandle(pflag));
  if (rank == 0) { int stt; waitpid(pid, &stt, 0); printf("child exit %d\n", WEXITSTATUS(stt)); }
  return 0;
}
