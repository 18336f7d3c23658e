// le_fix_host.inl -- host-side enqueue of the USER-LE events (included by le_engine.cu).

// RanMars::RanMars (src/random_mars.cpp:29-68) restated in 24-bit integers; s[] is chronological
// (oldest first): the reference's u[97], u[96], ..., u[1].
static void ranmars_seed_host(int seed, RngDev *st) {
  int u[98];
  int ij = (seed - 1) / 30082;
  int kl = (seed - 1) - 30082 * ij;
  int i = (ij / 177) % 177 + 2;
  int j = ij % 177 + 2;
  int k = (kl / 169) % 178 + 1;
  int l = kl % 169;
  for (int ii = 1; ii <= 97; ii++) {
    int s = 0, t = 1 << 23;
    for (int jj = 1; jj <= 24; jj++) {
      const int m = ((i * j) % 179) * k % 179;
      i = j; j = k; k = m;
      l = (53 * l + 1) % 169;
      if ((l * m) % 64 >= 32) s += t;
      t >>= 1;
    }
    u[ii] = s;
  }
  for (int q = 0; q < 97; q++) st->s[q] = u[97 - q];
  st->head = 0;
  st->c = 362436;
  st->consumed = 0;
}

// RanMars::uniform (src/random_mars.cpp:81-95), n times, results discarded
static void ranmars_skip_host(RngDev *st, long long n) {
  for (long long q = 0; q < n; q++) {
    int raw = st->s[st->head] - st->s[(st->head + 64) % 97];
    if (raw < 0) raw += 16777216;
    st->s[st->head] = raw;
    st->head = (st->head + 1) % 97;
    st->c -= 7654321;
    if (st->c < 0) st->c += 16777213;
  }
  st->consumed += n;
}

static int rng_index(int which) { return which == LE_FIX_EXTRUSION ? 0 : which == LE_FIX_EX_UNLOAD ? 1 : which == LE_FIX_EX_LOAD ? 2 : -1; }

static int rng_upload(le_ctx *c, int idx, int seed, long long consumed) {
  if (seed <= 0 || seed > 900000000) return fail(c, LE_EINVAL, "Invalid seed for Marsaglia random # generator");
  RngDev st;
  ranmars_seed_host(seed, &st);
  ranmars_skip_host(&st, 1);          // the constructor calls uniform() once (src/random_mars.cpp:67)
  st.consumed = 0;
  ranmars_skip_host(&st, consumed);
  CK(cudaMemcpyAsync(c->lf.rngdev + idx, &st, sizeof(RngDev), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->lf.rng[idx].seeded = 1; c->lf.rng[idx].seed = seed;
  return LE_OK;
}

static int ensure_rng(le_ctx *c, int idx, int seed) {
  if (c->lf.rng[idx].seeded) return LE_OK;
  return rng_upload(c, idx, seed, 0);
}

extern "C" int le_fix_rng_reset(le_ctx *c, int which, int seed, int64_t ndraws_consumed) {
  if (!c) return LE_EINVAL;
  const int idx = rng_index(which);
  if (idx < 0 || ndraws_consumed < 0) return fail(c, LE_EINVAL, "le_fix_rng_reset: bad arguments");
  if (!c->atoms_loaded) return fail(c, LE_ESTATE, "no atoms");
  cudaSetDevice(c->device);
  return rng_upload(c, idx, seed, ndraws_consumed);
}

extern "C" int le_fix_rng_consumed(le_ctx *c, int which, int64_t *ndraws) {
  if (!c || !ndraws) return LE_EINVAL;
  const int idx = rng_index(which);
  if (idx < 0) return fail(c, LE_EINVAL, "unknown fix");
  if (!c->atoms_loaded || !c->lf.rng[idx].seeded) { *ndraws = 0; return LE_OK; }
  cudaSetDevice(c->device);
  RngDev st;
  CK(cudaMemcpyAsync(&st, c->lf.rngdev + idx, sizeof(RngDev), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  *ndraws = st.consumed;
  return LE_OK;
}

/* the generator of a fix exactly as the reference holds it: RanMars::get_state / set_state (src/random_mars.cpp:297-319),
 * state[0..97] = u[0..97], [98] = i97, [99] = j97, [100] = c, [101] = cd, [102] = cm.  u[k] and c are multiples of 2^-24;
 * the ring s[q] = u[97 - q] * 2^24 never moves, head = 97 - i97. */
extern "C" int le_fix_rng_set_state(le_ctx *c, int which, const double *state) {
  if (!c || !state) return LE_EINVAL;
  const int idx = rng_index(which);
  if (idx < 0) return fail(c, LE_EINVAL, "unknown fix");
  if (!c->atoms_loaded) return fail(c, LE_ESTATE, "no atoms");
  const int i97 = (int)state[98], j97 = (int)state[99];
  int jexp = i97 - 64; if (jexp < 1) jexp += 97;
  if (i97 < 1 || i97 > 97 || j97 != jexp) return fail(c, LE_EINVAL, "le_fix_rng_set_state: not a RanMars state (i97 %d j97 %d)", i97, j97);
  RngDev st;
  for (int q = 0; q < 97; q++) {
    const double v = state[97 - q] * 16777216.0;
    st.s[q] = (int)llround(v);
    if (fabs(v - st.s[q]) > 1e-6 || st.s[q] < 0 || st.s[q] >= 16777216) return fail(c, LE_EINVAL, "le_fix_rng_set_state: u[%d] is not a 24-bit fraction", 97 - q);
  }
  st.head = 97 - i97;
  st.c = (int)llround(state[100] * 16777216.0);
  st.consumed = 0;
  cudaSetDevice(c->device);
  CK(cudaMemcpyAsync(c->lf.rngdev + idx, &st, sizeof(RngDev), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->lf.rng[idx].seeded = 1; c->lf.rng[idx].seed = 0;
  return LE_OK;
}

extern "C" int le_fix_rng_get_state(le_ctx *c, int which, double *state) {
  if (!c || !state) return LE_EINVAL;
  const int idx = rng_index(which);
  if (idx < 0) return fail(c, LE_EINVAL, "unknown fix");
  if (!c->atoms_loaded || !c->lf.rng[idx].seeded) return fail(c, LE_ESTATE, "the fix has not drawn yet and no state was set");
  cudaSetDevice(c->device);
  RngDev st;
  CK(cudaMemcpyAsync(&st, c->lf.rngdev + idx, sizeof(RngDev), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  state[0] = 0.0;
  for (int q = 0; q < 97; q++) state[97 - q] = st.s[q] / 16777216.0;
  const int i97 = 97 - st.head;
  int j97 = i97 - 64; if (j97 < 1) j97 += 97;
  state[98] = i97; state[99] = j97;
  state[100] = st.c / 16777216.0; state[101] = 7654321.0 / 16777216.0; state[102] = 16777213.0 / 16777216.0;
  return LE_OK;
}

// n (device-side count) draws of fix `idx`'s Marsaglia stream into lf.draws
static void ranmars_fill(le_ctx *c, int idx, const int *n_ptr) {
  LeFixDev &f = c->lf;
  const int nchunks = (f.draws_cap + RM_CHUNK - 1) / RM_CHUNK;
  LAUNCH(c, k_ranmars_base, 1, 32, (const RngDev *)(f.rngdev + idx), f.rm_base, n_ptr, f.draws_cap, c->d.ctrl);
  LAUNCH(c, k_ranmars_chunks, (nchunks + 3) / 4, 128, f.rngdev + idx, (const int *)f.rm_base, (const unsigned *)f.rm_jump, f.draws, n_ptr, f.draws_cap);
}

static void iscan(le_ctx *c, const int *in, int *out, int n, int *total) {
  const int nb = (n + 1023) / 1024;
  LAUNCH(c, k_iscan_partial, nb, 1024, in, n, c->lf.blocksum);
  LAUNCH(c, k_iscan_blocks, 1, 1024, c->lf.blocksum, nb, total);
  LAUNCH(c, k_iscan_apply, nb, 1024, in, out, n, (const int *)c->lf.blocksum);
}

static void compact_tasks(le_ctx *c) {
  LeFixDev &f = c->lf;
  iscan(c, f.flag, f.scan, c->N, f.counters + CNT_NTASK);
  LAUNCH(c, k_compact, grid_for(c->N, 256), 256, (const int *)f.flag, (const int *)f.scan, c->N, f.tasks);
}

static void draw_for_flagged(le_ctx *c, int idx, double fraction) {
  LeFixDev &f = c->lf;
  iscan(c, f.flag, f.scan, c->N, f.counters + CNT_NDRAW);
  ranmars_fill(c, idx, (const int *)(f.counters + CNT_NDRAW));
  LAUNCH(c, k_le_assign_draws, grid_for(c->N, 256), 256, f, c->N, (const int *)f.scan, fraction);
}

// ordered executor (le_fix.cuh): LE_EXEC_ROUNDS grid-wide rounds, then one block for whatever is left
template <class Task>
static void run_executor(le_ctx *c, const Task &T, unsigned base) {
  LeFixDev &f = c->lf;
  cudaMemsetAsync(f.exec_rem, 0, 64 * sizeof(int), c->stream);
  const int grid = std::min(grid_for(c->N, 256), 592);
  for (int round = 0; round < LE_EXEC_ROUNDS; round++) {
    LAUNCH(c, k_exec_claim<Task>, grid, 256, T, f.claim, f.done, (const int *)f.exec_rem, round, base);
    LAUNCH(c, k_exec_run<Task>, grid, 256, T, f.claim, f.done, f.exec_rem, round, base);
  }
  LAUNCH(c, k_exec_finish<Task>, 1, LE_EXEC_THREADS, T, f.claim, f.done, (const int *)f.exec_rem, base);
}

// update_topology sweep: detect the influenced atoms (around the marked end points), rebuild their special lists
static void topo_sweep(le_ctx *c, const int *marks, int mode) {
  LeFixDev &f = c->lf;
  int *nlist = f.counters + CNT_NLIST;
  const int *gate = (const int *)(f.counters + CNT_TOTAL);
  LAUNCH(c, k_le_topo_reset, 1, 1, nlist, f.mark_n);
  LAUNCH(c, k_le_topo_detect, 296, 256, c->d, f, marks, mode, gate, f.tasks, nlist);
  LAUNCH(c, k_le_topo_rebuild, 296, 128, c->d, (const int *)f.tasks, nlist);
}

static int enqueue_extrusion(le_ctx *c) {
  int r = ensure_rng(c, 0, c->fx.seed); if (r) return r;
  LeFixDev &f = c->lf;
  LeView V{c->d, f};
  ExtrusionArgs A{c->fx.btype, c->fx.neutral, c->fx.left, c->fx.right, c->fx.roadblock, c->fx.p};
  const int g = grid_for(c->N, 256);
  const int go = grid_for(c->d.gr0 - c->d.own0, 256);
  LAUNCH(c, k_le_begin, 1, 1, c->d);
  LAUNCH(c, k_ext_init, g, 256, V, A.btype);
  LAUNCH(c, k_ext_geo, go, 256, V, A);
  LAUNCH(c, k_le_exchange, 1, 1, c->d);
  LAUNCH(c, k_ext_visits, g, 256, V, A);
  compact_tasks(c);
  iscan(c, f.ndraw, f.scan2, c->N, f.counters + CNT_NDRAW);
  ranmars_fill(c, 0, (const int *)(f.counters + CNT_NDRAW));
  run_executor(c, VisitTask{V, A, (const int *)(f.counters + CNT_NTASK)}, 0u);
  LAUNCH(c, k_ext_flag, g, 256, V, 0);
  compact_tasks(c);
  run_executor(c, ReconcileTask{V, (const int *)(f.counters + CNT_NTASK)}, LE_EXEC_BASE_STRIDE);
  LAUNCH(c, k_ext_flag, g, 256, V, 1);
  compact_tasks(c);
  run_executor(c, BreakTask{V, (const int *)(f.counters + CNT_NTASK)}, 2u * LE_EXEC_BASE_STRIDE);
  LAUNCH(c, k_ext_create, g, 256, V, A.btype);
  LAUNCH(c, k_le_finish, 1, 1, c->d, f, 1);
  topo_sweep(c, (const int *)f.final_remove, 0);
  topo_sweep(c, (const int *)f.final_add, 1);
  return LE_OK;
}

static int enqueue_unload(le_ctx *c) {
  int r = ensure_rng(c, 1, c->fu.seed); if (r) return r;
  LeFixDev &f = c->lf;
  LeView V{c->d, f};
  UnloadArgs A{c->fu.btype, c->fu.rc * c->fu.rc, c->fu.prob};
  const int g = grid_for(c->N, 256);
  LAUNCH(c, k_le_begin, 1, 1, c->d);
  LAUNCH(c, k_unl_geo, grid_for(c->d.gr0 - c->d.own0, 256), 256, V, A);
  LAUNCH(c, k_le_exchange, 1, 1, c->d);
  LAUNCH(c, k_unl_candidates, g, 256, V, A);
  if (A.fraction < 1.0) draw_for_flagged(c, 1, A.fraction);
  LAUNCH(c, k_unl_break, g, 256, V, A);
  LAUNCH(c, k_le_finish, 1, 1, c->d, f, 2);
  topo_sweep(c, (const int *)f.final_remove, 0);
  return LE_OK;
}

static int enqueue_load(le_ctx *c) {
  int r = ensure_rng(c, 2, c->fl.seed); if (r) return r;
  // FixExLoad::init (fix_ex_load.cpp:217-218)
  {
    const int k = (c->fl.itype - 1) * c->ntypes + (c->fl.jtype - 1);
    if (!c->pair_set || c->fl.rc * c->fl.rc > c->cut[k] * c->cut[k])
      return fail(c, LE_EINVAL, c->fl.create ? "Fix bond/create cutoff is longer than pairwise cutoff" : "Fix ex_load cutoff is longer than pairwise cutoff");
  }
  LeFixDev &f = c->lf;
  LeView V{c->d, f};
  LoadArgs A{c->fl.btype, c->fl.itype, c->fl.jtype, c->fl.imax, c->fl.inew, c->fl.jmax, c->fl.jnew, c->fl.rc * c->fl.rc, c->fl.prob};
  const int g = grid_for(c->N, 256);
  LAUNCH(c, k_le_begin, 1, 1, c->d);
  if (c->fl.create) {
    // fix bond/create: bond counts from the first run's setup (enqueue_bond_create_setup), partner = closest eligible neighbor
    cudaMemcpyAsync(f.bondcount, f.bc_keep, sizeof(int) * c->N, cudaMemcpyDeviceToDevice, c->stream);
    LAUNCH(c, k_load_init, g, 256, V, A.btype, 0);
    LAUNCH(c, k_bcreate_geo, std::min(grid_for(c->d.gr0 - c->d.own0, 128), c->sm_count * 16), 128, V, A);
    LAUNCH(c, k_le_exchange, 1, 1, c->d);
    LAUNCH(c, k_bcreate_collect, g, 256, V);
  } else {
    LAUNCH(c, k_load_init, g, 256, V, A.btype, 1);
    LAUNCH(c, k_load_geo, grid_for(c->d.gr0 - c->d.own0, 256), 256, V, A);
    LAUNCH(c, k_le_exchange, 1, 1, c->d);
    LAUNCH(c, k_load_eligible, g, 256, V);
    LAUNCH(c, k_load_scan_runs, g, 256, V);
  }
  LAUNCH(c, k_load_flag_partners, g, 256, f, c->N);
  if (A.fraction < 1.0) draw_for_flagged(c, 2, A.fraction);
  LAUNCH(c, k_load_create, g, 256, V, A);
  if (c->fl.create) cudaMemcpyAsync(f.bc_keep, f.bondcount, sizeof(int) * c->N, cudaMemcpyDeviceToDevice, c->stream);
  LAUNCH(c, k_le_finish, 1, 1, c->d, f, 3);
  topo_sweep(c, (const int *)f.final_add, 1);
  return LE_OK;
}

// FixBondCreate::setup (fix_bond_create.cpp:302-345): the bonds of the fix's type are counted once, when its first run starts
static void enqueue_bond_create_setup(le_ctx *c) {
  if (!c->fl.on || !c->fl.create || c->fl.counted) return;
  LeView V{c->d, c->lf};
  LAUNCH(c, k_load_init, grid_for(c->N, 256), 256, V, c->fl.btype, 1);
  cudaMemcpyAsync(c->lf.bc_keep, c->lf.bondcount, sizeof(int) * c->N, cudaMemcpyDeviceToDevice, c->stream);
  c->fl.counted = 1;
}

// Modify::post_integrate on timestep `step`: each fix checks its own gate
// (fix_extrusion.cpp:265 `ntimestep % nevery - 1`, fix_ex_unload.cpp:178 `- 2`, fix_ex_load.cpp:338 `- 3`)
static int enqueue_le_events(le_ctx *c, int64_t step) {
  bool any = false;
  for (int which : c->fix_order) {
    int r = LE_OK;
    const bool fires = (which == LE_FIX_EXTRUSION && c->fx.on && (step % c->fx.nevery - 1) == 0) ||
                       (which == LE_FIX_EX_UNLOAD && c->fu.on && (step % c->fu.nevery - c->fu.phase) == 0) ||
                       (which == LE_FIX_EX_LOAD && c->fl.on && (step % c->fl.nevery - c->fl.phase) == 0);
    if (fires && any && c->nranks > 1) {
      // a second event on the same timestep (fix bond/create and fix bond/break share their steps): its owners store their
      // geometry records into the same per-tag slots of every GPU, so nobody may start before every GPU has read the records of
      // the event before -- one more flag round (events on different timesteps are separated by the per-step rounds)
      LAUNCH(c, k_le_begin, 1, 1, c->d);
      LAUNCH(c, k_le_exchange, 1, 1, c->d);
    }
    if (which == LE_FIX_EXTRUSION && c->fx.on && (step % c->fx.nevery - 1) == 0) { r = enqueue_extrusion(c); any = true; }
    else if (which == LE_FIX_EX_UNLOAD && c->fu.on && (step % c->fu.nevery - c->fu.phase) == 0) { r = enqueue_unload(c); any = true; }
    else if (which == LE_FIX_EX_LOAD && c->fl.on && (step % c->fl.nevery - c->fl.phase) == 0) { r = enqueue_load(c); any = true; }
    if (r) return r;
  }
  // the digests of the atoms an event touched were refreshed by its topology sweeps (k_le_topo_rebuild)
  return any ? LE_OK : LE_OK;
}

extern "C" int le_run_le_event(le_ctx *c, int which) {
  if (!c) return LE_EINVAL;
  int r = ensure_ready(c); if (r) return r;
  if (!c->lists_valid) return fail(c, LE_ESTATE, "no neighbor/bond lists: call le_force_rebuild or le_run first");
  if (which == LE_FIX_EXTRUSION) { if (!c->fx.on) return fail(c, LE_ESTATE, "fix extrusion not defined"); r = enqueue_extrusion(c); }
  else if (which == LE_FIX_EX_UNLOAD) { if (!c->fu.on) return fail(c, LE_ESTATE, "fix ex_unload not defined"); r = enqueue_unload(c); }
  else if (which == LE_FIX_EX_LOAD) { if (!c->fl.on) return fail(c, LE_ESTATE, "fix ex_load not defined"); enqueue_bond_create_setup(c); r = enqueue_load(c); }
  else return fail(c, LE_EINVAL, "unknown fix");
  if (r) return r;
  return sync_and_check(c);
}
