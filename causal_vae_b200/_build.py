"""In-tree build of libcvae_b200.so (nvcc, sm_100a only).  Cross-compiles without a GPU."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcvae_b200.so")
SOURCES = ["conv.cu", "conv_tc.cu", "conv_halo_tc.cu", "wgrad_tc.cu", "wgrad_tile.cu", "skinny.cu", "conv_few.cu", "pack_batch.cu", "linear_small.cu", "head_bwd.cu", "norm.cu", "attention.cu", "elementwise.cu", "loss_optim.cu", "classify.cu", "input_pipeline.cu", "attention_long.cu", "graph_prio.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "cvae_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, timing=False):
    """timing=True: the role-timer build (-DCVAE_TIMING) as libcvae_b200_timing.so beside the product library
    (loaded only through CVAE_LIB by scripts/bench_layers.py)."""
    timing = timing or os.environ.get("CVAE_TIMING") == "1"
    lib_out = LIB.replace(".so", "_timing.so") if timing else LIB
    if not force and not timing and not needs_build():
        return LIB
    objs = []
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    if timing:       # role-level wait accounting in the tensor-core kernels (debug)
        flags.append("-DCVAE_TIMING")
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", "_t.o" if timing else ".o"))
        cmd = [_nvcc(), *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [_nvcc(), "-shared", "-o", lib_out, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    subprocess.check_call(cmd)
    return lib_out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, timing="--timing" in sys.argv))
