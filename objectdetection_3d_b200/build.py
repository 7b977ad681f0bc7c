"""Build libpp_b200.so (sm_100a only) in-tree with nvcc.

    python -m objectdetection_3d_b200.build [--force]

The shared library lands next to this file so that it travels with the repository snapshot;
it is git-ignored (source-only history).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpp_b200.so")
SOURCES = ["pp_api.cu", "pp_sort.cu", "pp_voxelize.cu", "pp_pillar.cu", "pp_boxes.cu", "pp_nms.cu", "pp_points.cu", "pp_topk.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",            # every FP op rounds like the eager reference op it mirrors
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--cudart", "static",
]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "pp_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    extra = os.environ.get("PP_NVCC_EXTRA", "").split()      # development only (e.g. -DPP_TIMING): a separate library
    lib = LIB
    if extra:
        force = True
        lib = os.path.join(HERE, "libpp_b200_dev.so")
    if not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, "build_dev" if extra else "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed on %s" % src)
    cmd = [nvcc(), "-shared", "--cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
