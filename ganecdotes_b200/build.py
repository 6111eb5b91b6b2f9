"""In-tree build of the CUDA library (sm_100a) with nvcc.  No torch dependency:
the library is a plain C-ABI shared object (include/ganecdotes_b200.h)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libganecdotes_b200.so")
SOURCES = ["gx_api.cu", "gx_umma.cu", "gx_fir.cu", "gx_synthesis.cu", "gx_head.cu", "gx_exchange.cu", "gx_index.cu", "gx_simclr.cu", "gx_conv_generic.cu", "gx_kmeans.cu", "gx_conv_small.cu"]
HEADERS = ["gx_common.cuh", "gx_ptx.cuh", "gx_ll.cuh", os.path.join("..", "..", "include", "ganecdotes_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    for f in SOURCES + HEADERS:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into csrc/libganecdotes_b200.so."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH] + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
