"""Builds smmd/_C (the PyTorch C++ extension wrapper of libsmmd.so) in-tree; called by __graft_entry__.build().
The extension contains no device code: it is host C++ that links against ../../lib/libsmmd.so."""
import glob
import os
import shutil
import sys


def build(verbose=False):
    from torch.utils import cpp_extension

    here = os.path.dirname(os.path.abspath(__file__))
    pkg = os.path.dirname(os.path.dirname(here))
    libdir = os.path.join(pkg, "lib")
    bdir = os.path.join(pkg, "build", "torch_ext")
    os.makedirs(bdir, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    try:
        cpp_extension.load(name="_C", sources=[os.path.join(here, "smmd_torch.cpp")], build_directory=bdir,
                           extra_cflags=["-O2", "-std=c++17"], with_cuda=True,
                           extra_ldflags=["-L" + libdir, "-lsmmd", "-Wl,-rpath,\\$$ORIGIN/../lib"], is_python_module=False,
                           verbose=verbose)
    except OSError:
        pass   # compiled and linked; loading it from the build directory fails (its rpath is relative to smmd/)
    built = glob.glob(os.path.join(bdir, "_C*.so"))
    if not built:
        raise RuntimeError("smmd._C was not produced in %s" % bdir)
    dst = os.path.join(pkg, "smmd", "_C.so")
    shutil.copyfile(built[0], dst)
    return dst


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
