"""Development: build an instrumented variant of libdwt_b200.so next to this file (one source recompiled with an extra
-D flag, the other objects reused from dwt_b200/lib) and return its path.  Needs the normal library to be built first."""
import importlib.util, os, subprocess
HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(HERE, "..", "..", "..", "dwt-domain-adaptation_b200", "dwt_b200")


def build(name, define, source):
    spec = importlib.util.spec_from_file_location("dwt_b200_build", os.path.join(PKG, "build.py"))
    B = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(B)                          # dwt_b200/build.py, without putting the package dir on sys.path
    B.build()
    out = os.path.join(HERE, f"libdwt_b200_{name}.so")
    src = os.path.join(B.CSRC, source)
    if os.path.exists(out) and os.path.getmtime(out) > max(os.path.getmtime(src), os.path.getmtime(B.LIB)):
        return out
    obj = os.path.join(HERE, f"{source[:-3]}_{name}.o")
    subprocess.check_call([B._nvcc(), *B.NVCC_FLAGS, f"-D{define}", "-c", src, "-o", obj])
    others = [os.path.join(PKG, "lib", s.replace(".cu", ".o")) for s in B.SOURCES if s != source]
    subprocess.check_call([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", *others, obj, "-o", out])
    os.remove(obj)
    return out
