"""Build recipe for libydst.so (sm_100a only).  `python -m yolo_deepsort_b200.build` or __graft_entry__.build().

nvcc cross-compiles without a GPU; the .so is built IN-TREE (yolo_deepsort_b200/libydst.so) so it travels to
the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", f) for f in ("api.cu", "conv_tc.cu", "layers.cu", "nms.cu", "assoc.cu", "cosine_tc.cu", "net.cu", "tracker.cu", "action.cu")]
OUT = os.path.join(HERE, "libydst.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-ffp-contract=off", "--fmad=true", "-cudart", "shared", "-Xptxas", "-v"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))] + [os.path.join(HERE, "..", "include", "ydst.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = [NVCC] + FLAGS + SRC + ["-o", OUT]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libydst.so")
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if verbose:
        print(log)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
