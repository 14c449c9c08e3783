"""Recipe for oracle/_ref (TEST INFRASTRUCTURE): the reference DCNv2 on the GPU.

Compiles the reference's two THC-free kernel files UNMODIFIED, from where they lie under
/root/reference, for sm_100a, together with oracle/ref_host.cu (the cuBLAS restatement of the
removed-THC host loop, dcn_v2_cuda.c:61-97,161-231,243-330) into

    oracle/_ref/libdcnv2_ref.so

No reference source is copied into the repository: only the built library lands in oracle/_ref/
(git-ignored, NOT gpurun-ignored, so it travels to the GPU box like our own .so files).  On the
GPU box /root/reference does not exist; build() then only reports whether the prebuilt library
is there.  Run: ``python -m oracle.build_ref [--force]``.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "libdcnv2_ref.so")
REF_CUDA = "/root/reference/src/lib/models/networks/DCNv2/src/cuda"
REF_SOURCES = [os.path.join(REF_CUDA, "dcn_v2_im2col_cuda.cu"),
               os.path.join(REF_CUDA, "dcn_v2_psroi_pooling_cuda.cu")]
HOST = os.path.join(HERE, "ref_host.cu")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def available():
    return os.path.exists(LIB)


def build(force=False):
    """Returns the library path, or None when neither the reference sources nor a prebuilt
    library are present (never raises for a missing reference: the checker is optional)."""
    have_src = all(os.path.exists(s) for s in REF_SOURCES)
    if not have_src:
        return LIB if available() else None
    srcs = REF_SOURCES + [HOST]
    if not force and available() and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in srcs):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-shared", "-Xcompiler", "-fPIC",
           "-cudart", "static", "-I" + REF_CUDA] + srcs + ["-lcublas", "-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
