"""Build libmst_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

    python new-vit_b200/build.py            # or: python -c "import __graft_entry__ as g; g.build()"

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmst_b200.so")
SOURCES = ["api.cu", "gemm_tc.cu", "gemm_wt.cu", "attention_tc16.cu", "attention_tcg.cu", "attention_mma.cu", "kernels.cu", "extras.cu", "prep.cu", "train.cu", "train_enc.cu", "gemm_wgrad.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build_library(force=False, verbose=False, experiments=False, out=None):
    """experiments=True adds -DMST_EXPERIMENTS: the MST_* environment switches (alternative tilings, A/B toggles) and the
    clock64() phase counters behind mst_debug_gemm_timing.  The product build has neither."""
    global LIB
    lib = out or LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "mst_b200.h"))
    if not force and os.path.exists(lib) and os.path.getmtime(lib) >= _newest(deps):
        return lib
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    bdir = os.path.join(HERE, "build", "exp" if experiments else "prod")
    os.makedirs(bdir, exist_ok=True)
    for s in srcs:
        o = os.path.join(bdir, os.path.basename(s) + ".o")
        objs.append(o)
        cmd = [nvcc] + NVCC_FLAGS + (["-DMST_EXPERIMENTS"] if experiments else []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            print(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [nvcc, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    exp = "--experiments" in sys.argv   # the experiments build goes to its own file; select it with MST_LIB_PATH
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, experiments=exp,
                        out=os.path.join(HERE, "libmst_b200_exp.so") if exp else None))
