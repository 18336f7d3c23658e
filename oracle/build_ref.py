#!/usr/bin/env python3
"""Recipe that compiles the reference's own CPU implementation into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported, linked or executed
by the product path (lammps_le_b200/); only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may run what this builds.

What it does (no cmake, none of the reference's own build scripts are run):
  * compiles the reference sources WHERE THEY LIE under /root/reference/src
    (core src/*.cpp + MOLECULE/*.cpp + USER-LE/*.cpp + STUBS/mpi.c) with g++ -O2,
    object files and outputs only into oracle/_ref/;
  * writes the small style_*.h include lists LAMMPS expects (they are plain
    lists of `#include "x.h"` lines selected by the *_CLASS marker each style
    header carries; the reference makes them with src/Make.sh:18-54 or
    cmake/Modules/StyleHeaderUtils.cmake) into oracle/_ref/gen/;
  * links oracle/_ref/liblammps_ref.so, oracle/_ref/lmp_ref (src/main.cpp) and
    oracle/_ref/ref_harness (oracle/ref_harness.cpp, our own state-capture driver).

No reference source is copied into this repository.  oracle/_ref/ is git-ignored
but travels to the GPU box with the gpurun snapshot.
"""
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LE_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
GEN = os.path.join(OUT, "gen")
OBJ = os.path.join(OUT, "obj")

SRC_DIRS = ["src", "src/MOLECULE", "src/USER-LE"]
# single styles from packages that are not compiled as a whole: fix bond/break and fix bond/create (src/MC), the ancestors of fix
# ex_unload and fix ex_load -- the
# checkers for le_fix_bond_break and le_fix_bond_create (tests/test_gpu_md.py)
EXTRA_STYLES = [("src/MC", "fix_bond_break"), ("src/MC", "fix_bond_create")]
CXXFLAGS = ["-O2", "-std=c++11", "-fPIC", "-DLAMMPS_SMALLBIG", "-DLAMMPS_EXCEPTIONS",
            "-ffp-contract=off", "-w"]
OMP = False      # second build (oracle/_ref/omp/): the same sources + the USER-OMP styles whose base style is compiled, -fopenmp


def set_variant(omp):
    """switch the module-level paths / flags between the serial build and the threaded one (bench.py's reference arm)"""
    global OUT, GEN, OBJ, SRC_DIRS, CXXFLAGS, OMP
    OMP = omp
    OUT = os.path.join(HERE, "_ref", "omp") if omp else os.path.join(HERE, "_ref")
    GEN, OBJ = os.path.join(OUT, "gen"), os.path.join(OUT, "obj")
    SRC_DIRS = ["src", "src/MOLECULE", "src/USER-LE"] + (["src/USER-OMP"] if omp else [])
    # the serial build is the bit-exact checker the golden vectors were made with (-O2, no contraction); the threaded
    # build is only ever TIMED (bench.py's reference arm / cpu_baseline): -O3 as the reference's own cmake Release build
    CXXFLAGS = ["-O3" if omp else "-O2", "-std=c++11", "-fPIC", "-DLAMMPS_SMALLBIG", "-DLAMMPS_EXCEPTIONS", "-ffp-contract=off", "-w"]
    if omp:
        CXXFLAGS += ["-fopenmp", "-DLMP_USER_OMP"]


def omp_file_wanted(name):
    """src/USER-OMP/Install.sh: x_omp.{cpp,h} is installed only if x.{cpp,h} exists among the compiled sources"""
    if name in ("thr_omp.cpp", "thr_omp.h", "thr_data.cpp", "thr_data.h"):
        return True
    m = re.match(r"(.*)_omp\.(cpp|h)$", name)
    if not m:
        return False
    base = "%s.%s" % (m.group(1), m.group(2))
    return any(os.path.exists(os.path.join(REF, d, base)) for d in ("src", "src/MOLECULE", "src/USER-LE"))

# (marker in header, filename prefix regex, style_<name>.h)
STYLES = [
    ("ANGLE_CLASS", r"angle_", "angle"), ("ATOM_CLASS", r"atom_vec_", "atom"),
    ("BODY_CLASS", r"body_", "body"), ("BOND_CLASS", r"bond_", "bond"),
    ("COMMAND_CLASS", r"", "command"), ("COMPUTE_CLASS", r"compute_", "compute"),
    ("DIHEDRAL_CLASS", r"dihedral_", "dihedral"), ("DUMP_CLASS", r"dump_", "dump"),
    ("FIX_CLASS", r"fix_", "fix"), ("IMPROPER_CLASS", r"improper_", "improper"),
    ("INTEGRATE_CLASS", r"", "integrate"), ("KSPACE_CLASS", r"", "kspace"),
    ("MINIMIZE_CLASS", r"min_", "minimize"), ("NBIN_CLASS", r"nbin_", "nbin"),
    ("NPAIR_CLASS", r"npair_", "npair"), ("NSTENCIL_CLASS", r"nstencil_", "nstencil"),
    ("NTOPO_CLASS", r"ntopo_", "ntopo"), ("PAIR_CLASS", r"pair_", "pair"),
    ("READER_CLASS", r"reader_", "reader"), ("REGION_CLASS", r"region_", "region"),
]


def up_to_date(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps if os.path.exists(d))


CHANGED_STYLE_LISTS = []


def write_if_changed(path, text):
    if os.path.exists(path) and open(path).read() == text:
        return
    with open(path, "w") as f:
        f.write(text)
    CHANGED_STYLE_LISTS.append(os.path.basename(path))


def gen_headers():
    os.makedirs(GEN, exist_ok=True)
    headers = []
    for d in SRC_DIRS:
        p = os.path.join(REF, d)
        headers += [os.path.join(p, h) for h in sorted(os.listdir(p)) if h.endswith(".h") and (d != "src/USER-OMP" or omp_file_wanted(h))]
    headers += [os.path.join(REF, d, b + ".h") for d, b in EXTRA_STYLES]
    first_lines = {}
    for h in headers:
        with open(h, errors="replace") as f:
            first_lines[h] = f.read()
    for marker, prefix, name in STYLES:
        names = []
        for h in headers:
            base = os.path.basename(h)
            if prefix and not base.startswith(prefix):
                continue
            if re.search(r"^#ifdef %s" % marker, first_lines[h], re.M):
                names.append(base)
        write_if_changed(os.path.join(GEN, "style_%s.h" % name),
                         "".join('#include "%s"\n' % n for n in sorted(set(names))))
        # packages_<kind>.h only feeds "which package provides style X" hints in
        # error messages (src/lammps.cpp:909-1010); an empty list is valid.
        write_if_changed(os.path.join(GEN, "packages_%s.h" % name), "")
    write_if_changed(os.path.join(GEN, "lmpinstalledpkgs.h"),
                     "#ifndef LMP_INSTALLED_PKGS_H\n#define LMP_INSTALLED_PKGS_H\n"
                     "const char * LAMMPS_NS::LAMMPS::installed_packages[] = "
                     '{"MOLECULE", "USER-LE"%s, NULL};\n#endif\n' % (', "USER-OMP"' if OMP else ""))
    write_if_changed(os.path.join(GEN, "lmpgitversion.h"),
                     "#ifndef LMP_GIT_VERSION_H\n#define LMP_GIT_VERSION_H\n"
                     "const bool LAMMPS_NS::LAMMPS::has_git_info = false;\n"
                     'const char LAMMPS_NS::LAMMPS::git_commit[] = "(unknown)";\n'
                     'const char LAMMPS_NS::LAMMPS::git_branch[] = "(unknown)";\n'
                     'const char LAMMPS_NS::LAMMPS::git_descriptor[] = "(unknown)";\n#endif\n')


def includes():
    inc = ["-I" + GEN, "-I" + os.path.join(REF, "src/STUBS")]
    inc += ["-I" + os.path.join(REF, d) for d in SRC_DIRS]
    inc += ["-I" + os.path.join(REF, d) for d in sorted({d for d, _ in EXTRA_STYLES})]
    return inc


def compile_one(job):
    src, obj, cc = job
    if up_to_date(obj, [src]):
        return 0, src, ""
    cmd = [cc] + (CXXFLAGS if cc == "g++" else ["-O2", "-fPIC", "-w"]) + includes() + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return r.returncode, src, r.stderr


def build():
    if not os.path.isdir(os.path.join(REF, "src")):
        if os.path.exists(os.path.join(OUT, "lmp_ref" if OMP else "ref_harness")):
            print("oracle/_ref: reference tree absent, using prebuilt files")
            return 0
        print("oracle/_ref: reference tree absent and nothing prebuilt", file=sys.stderr)
        return 1
    os.makedirs(OBJ, exist_ok=True)
    del CHANGED_STYLE_LISTS[:]
    gen_headers()
    # the objects that include a style list do not notice its change through their own source's time stamp
    users = {"style_fix.h": ["modify.o", "lammps.o"], "style_integrate.h": ["update.o", "lammps.o"]}
    for lst in CHANGED_STYLE_LISTS:
        for o in users.get(lst, []):
            if os.path.exists(os.path.join(OBJ, o)):
                os.remove(os.path.join(OBJ, o))
    jobs = []
    for d in SRC_DIRS:
        p = os.path.join(REF, d)
        for s in sorted(os.listdir(p)):
            if s.endswith(".cpp") and s != "main.cpp" and (d != "src/USER-OMP" or omp_file_wanted(s)):
                jobs.append((os.path.join(p, s), os.path.join(OBJ, s[:-4] + ".o"), "g++"))
    for d, b in EXTRA_STYLES:
        jobs.append((os.path.join(REF, d, b + ".cpp"), os.path.join(OBJ, b + ".o"), "g++"))
    jobs.append((os.path.join(REF, "src/STUBS/mpi.c"), os.path.join(OBJ, "mpi_stubs.o"), "gcc"))
    nthreads = int(os.environ.get("LE_BUILD_JOBS", str(os.cpu_count() or 4)))
    failed = []
    with ThreadPoolExecutor(nthreads) as ex:
        for rc, src, err in ex.map(compile_one, jobs):
            if rc:
                failed.append((src, err))
    if failed:
        for src, err in failed:
            print("FAILED", src, "\n", err[-2000:], file=sys.stderr)
        return 1
    objs = [j[1] for j in jobs]
    lib = os.path.join(OUT, "liblammps_ref.so")
    if not up_to_date(lib, objs):
        subprocess.check_call(["g++", "-shared", "-o", lib] + objs + (["-fopenmp"] if OMP else []))
    rpath = "-Wl,-rpath,$ORIGIN"
    lmp = os.path.join(OUT, "lmp_ref")
    main_cpp = os.path.join(REF, "src/main.cpp")
    if not up_to_date(lmp, [lib, main_cpp]):
        subprocess.check_call(["g++"] + CXXFLAGS + includes() + [main_cpp, "-o", lmp, "-L" + OUT,
                                                                 "-llammps_ref", rpath])
    harness_src = os.path.join(HERE, "ref_harness.cpp")
    harness = os.path.join(OUT, "ref_harness")
    if not OMP and os.path.exists(harness_src) and not up_to_date(harness, [lib, harness_src]):
        subprocess.check_call(["g++"] + CXXFLAGS + includes() + [harness_src, "-o", harness, "-L" + OUT,
                                                                 "-llammps_ref", rpath])
    print("oracle/_ref built:", lib)
    return 0


def build_b200():
    """oracle/_ref/b200/lmp_b200: the serial reference objects + the `run_style le/b200` binding a maintainer would add
    (lammps_le_b200/lammps_style/verlet_le_b200.{h,cpp}), linked against lammps_le_b200/libleb200.so.  Only update.cpp and
    lammps.cpp are recompiled (the integrator table and the -h style list come from style_integrate.h).  TEST-ONLY: tests/test_gpu_lammps_style.py drives decks
    through the reference's own Input::file into the engine with it."""
    set_variant(False)
    root = os.path.dirname(HERE)
    style_dir = os.path.join(root, "lammps_le_b200", "lammps_style")
    lib = os.path.join(root, "lammps_le_b200", "libleb200.so")
    out = os.path.join(OUT, "b200")
    exe = os.path.join(out, "lmp_b200")
    if not os.path.isdir(os.path.join(REF, "src")):
        print("oracle/_ref/b200: reference tree absent, %s" % ("using the prebuilt binary" if os.path.exists(exe) else "nothing prebuilt"))
        return 0 if os.path.exists(exe) else 1
    if not os.path.exists(lib) or not os.path.isdir(OBJ):
        print("oracle/_ref/b200: needs libleb200.so and the serial reference objects first", file=sys.stderr)
        return 1
    gen = os.path.join(out, "gen")
    os.makedirs(gen, exist_ok=True)
    for f in os.listdir(GEN):
        text = open(os.path.join(GEN, f)).read()
        if f == "style_integrate.h":
            text += '#include "verlet_le_b200.h"\n'
        write_if_changed(os.path.join(gen, f), text)
    inc = ["-I" + gen, "-I" + style_dir, "-I" + os.path.join(root, "include"), "-I" + os.path.join(REF, "src/STUBS")] + ["-I" + os.path.join(REF, d) for d in SRC_DIRS]
    inc += ["-I" + os.path.join(REF, d) for d in sorted({d for d, _ in EXTRA_STYLES})]
    gen_files = [os.path.join(gen, f) for f in os.listdir(gen)]
    srcs = [(os.path.join(REF, "src/update.cpp"), os.path.join(out, "update.o")), (os.path.join(REF, "src/lammps.cpp"), os.path.join(out, "lammps.o")), (os.path.join(style_dir, "verlet_le_b200.cpp"), os.path.join(out, "verlet_le_b200.o"))]
    for src, obj in srcs:
        if not up_to_date(obj, [src, os.path.join(style_dir, "verlet_le_b200.h"), os.path.join(root, "include", "le_b200.h")] + gen_files):
            subprocess.check_call(["g++"] + CXXFLAGS + ["-DLE_B200_WITH_MC"] + inc + ["-c", src, "-o", obj])
    objs = [os.path.join(OBJ, o) for o in sorted(os.listdir(OBJ)) if o.endswith(".o") and o not in ("update.o", "lammps.o")] + [o for _, o in srcs]
    if not up_to_date(exe, objs + [lib]):
        subprocess.check_call(["g++"] + CXXFLAGS + inc + [os.path.join(REF, "src/main.cpp")] + objs + ["-o", exe, "-L" + os.path.dirname(lib), "-lleb200",
                                                                                                 "-Wl,-rpath,$ORIGIN/../../../lammps_le_b200"])
    print("oracle/_ref/b200 built:", exe)
    return 0


def main():
    set_variant(False)
    rc = build()
    if rc == 0:
        try:
            if build_b200():
                print("oracle/_ref/b200: run_style le/b200 build failed", file=sys.stderr)
        except Exception as ex:
            print("oracle/_ref/b200: %s" % ex, file=sys.stderr)
    if rc == 0 and os.path.isdir(os.path.join(REF, "src/USER-OMP")) and os.environ.get("LE_REF_OMP", "1") == "1":
        # the threaded build only feeds bench.py's reference arm; its failure must not take the oracle down
        set_variant(True)
        try:
            if build():
                print("oracle/_ref/omp: threaded reference build failed (bench.py falls back to the serial one)", file=sys.stderr)
        except Exception as ex:
            print("oracle/_ref/omp: %s" % ex, file=sys.stderr)
        set_variant(False)
    return rc


if __name__ == "__main__":
    sys.exit(main())
