"""Builds integration/_build/kaolin_b200_binding.so: INTEGRATION.md Option A (the reference-side
pybind11 / ATen host wrappers on top of the C ABI) compiled for real — g++ against the torch
headers, linked to kaolin_b200/csrc/libdibr_b200.so (rpath-relative).  No reference sources are
involved; only torch's headers and this repo."""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_build")
NAME = "kaolin_b200_binding"
SRC = os.path.join(HERE, "kaolin_binding.cpp")
LIBDIR = os.path.join(ROOT, "kaolin_b200", "csrc")


def build(force=False, verbose=True):
    so = os.path.join(OUT, NAME + ".so")
    deps = [SRC, os.path.join(ROOT, "include", "dibr_b200.h")]
    if not force and os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(d) for d in deps):
        return so
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT, exist_ok=True)
    inc = [f"-I{p}" for p in ce.include_paths("cuda")]
    inc += [f"-I{sysconfig.get_paths()['include']}", f"-I{os.path.join(ROOT, 'include')}"]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    libdirs = ce.library_paths("cuda")
    cmd = (["g++", "-O2", "-std=c++17", "-fPIC", "-shared", SRC, "-o", so, "-DWITH_CUDA",
            f"-DTORCH_EXTENSION_NAME={NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={abi}"]
           + inc + [f"-L{d}" for d in libdirs] + [f"-Wl,-rpath,{d}" for d in libdirs]
           + [f"-L{LIBDIR}", "-Wl,-rpath,$ORIGIN/../../kaolin_b200/csrc", "-ldibr_b200",
              "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"])
    if verbose:
        print("[build_binding]", " ".join(cmd[:8]), "...")
    subprocess.check_call(cmd)
    return so


def load():
    import importlib.util
    import torch  # noqa: F401
    so = os.path.join(OUT, NAME + ".so")
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
