"""Builds the reference's OWN native ops (lib/gan/optim/{upfirdn2d,fused_bias_act}{.cpp,_kernel.cu}) for sm_100a
from the sources where they lie under /root/reference, into oracle/_ref/ (git-ignored; travels to the GPU box with
the snapshot).  Test / benchmark infrastructure only: the GPU-side reference beside gx_upfirdn2d and
gx_fused_bias_act ("this SIMT kernel is the GPU baseline to beat").  Nothing is copied into the repo;
nothing under ganecdotes_b200/ loads these modules.

    python oracle/build_ref.py          # in the build container (nvcc cross-compiles, no GPU needed)

On the GPU box /root/reference does not exist: `load()` then imports the prebuilt modules, or returns None."""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("GX_REFERENCE", "/root/reference")
MODULES = {
    "ref_upfirdn2d": ["upfirdn2d.cpp", "upfirdn2d_kernel.cu"],
    "ref_fused_bias_act": ["fused_bias_act.cpp", "fused_bias_act_kernel.cu"],
}


def build(verbose=False):
    """JIT-free build with torch.utils.cpp_extension into oracle/_ref/<name>/<name>.so"""
    from torch.utils import cpp_extension
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    src_dir = os.path.join(REF, "lib", "gan", "optim")
    if not os.path.isdir(src_dir):
        return False
    for name, files in MODULES.items():
        bdir = os.path.join(OUT, name)
        os.makedirs(bdir, exist_ok=True)
        so = os.path.join(bdir, name + ".so")
        srcs = [os.path.join(src_dir, f) for f in files]
        if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
            continue
        cpp_extension.load(name, sources=srcs, build_directory=bdir, verbose=verbose, is_python_module=False,
                           extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"])
    return True


def load(name):
    """import a prebuilt reference module (None if it was not built)"""
    so = os.path.join(OUT, name, name + ".so")
    if not os.path.exists(so):
        return None
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location(name, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ok = build(verbose="-v" in sys.argv)
    print("built" if ok else "reference sources not found", OUT)
