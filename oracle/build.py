"""Build recipe for the checker side: the C restatement and, when the
reference tree is mounted, the reference's own sources compiled where they lie.

TEST INFRASTRUCTURE ONLY (see oracle/scda_oracle.c header).

Outputs
-------
oracle/liboracle.so            gcc -O2 -fopenmp -ffp-contract=off of scda_oracle.c
oracle/_ref/libscda_ref.so     the six reference .cu files, UNMODIFIED, nvcc sm_100a
                               (launcher symbols have C linkage: ROIPoolForwardLaucher,
                               ROIAlignForwardLaucher, _nms, IOUOverlap, *FocalLoss*Laucher)
oracle/_ref/cython_bbox*.so    the reference's cython_bbox.pyx, cythonized + gcc

`oracle/_ref/` is git-ignored and is only (re)built when /root/reference exists
(this container); the GPU box uses the prebuilt files that travel with the
snapshot.  No reference source is copied into the repo: the compilers read the
files under /root/reference and intermediate files go to a temp dir.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("SCDA_REFERENCE_ROOT", "/root/reference")
REF_OUT = os.path.join(HERE, "_ref")

REF_CU = [
    ("extensions/_roi_pooling/src", "roi_pooling_kernel.cu"),
    ("extensions/_roi_align/src", "roi_align_kernel.cu"),
    ("extensions/_nms/src/cuda", "nms_kernel.cu"),
    ("extensions/_bbox_helper/src/cuda", "iou_overlap_kernel.cu"),
    ("extensions/_focal_loss/src/cuda", "focal_loss_sigmoid_kernel.cu"),
    ("extensions/_focal_loss/src/cuda", "focal_loss_softmax_kernel.cu"),
]


def _newer(target: str, *sources: str) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def _run(cmd: list[str], **kw) -> None:
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))


def build_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "scda_oracle.c")
    out = os.path.join(HERE, "liboracle.so")
    if force or not _newer(out, src, __file__):
        _run(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fvisibility=hidden", "-shared",
              "-fPIC", "-o", out, src, "-lm"])
    return out


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "extensions"))


def build_ref_cuda(force: bool = False) -> str | None:
    """nvcc the reference's .cu files in place into oracle/_ref/libscda_ref.so."""
    out = os.path.join(REF_OUT, "libscda_ref.so")
    if not reference_available():
        return out if os.path.exists(out) else None
    srcs = [os.path.join(REF_ROOT, d, f) for d, f in REF_CU]
    if not force and _newer(out, *srcs, __file__):
        return out
    os.makedirs(REF_OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        objs = []
        for d, f in REF_CU:
            obj = os.path.join(tmp, f.replace(".cu", ".o"))
            _run(["nvcc", "-c", "-O2", "-x", "cu", "-Xcompiler", "-fPIC",
                  "-gencode", "arch=compute_100a,code=sm_100a",
                  "-I", os.path.join(REF_ROOT, d), "-o", obj, os.path.join(REF_ROOT, d, f)])
            objs.append(obj)
        _run(["nvcc", "-shared", "-o", out, "-gencode", "arch=compute_100a,code=sm_100a",
              "--cudart", "shared"] + objs)
    return out


def build_ref_cython(force: bool = False) -> str | None:
    """Cythonize + compile the reference's cython_bbox.pyx into oracle/_ref/."""
    ext_suffix = sysconfig.get_config_var("EXT_SUFFIX")
    out = os.path.join(REF_OUT, "cython_bbox" + ext_suffix)
    pyx = os.path.join(REF_ROOT, "extensions/_cython_bbox/cython_bbox.pyx")
    if not reference_available():
        return out if os.path.exists(out) else None
    if not force and _newer(out, pyx, __file__):
        return out
    import numpy as np
    os.makedirs(REF_OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        c_file = os.path.join(tmp, "cython_bbox.c")
        _run([sys.executable, "-m", "cython", "-3", "-o", c_file, pyx])
        _run(["gcc", "-O2", "-shared", "-fPIC", "-Wno-cpp", "-Wno-unused-function",
              "-I", sysconfig.get_paths()["include"], "-I", np.get_include(),
              "-o", out, c_file])
    return out


def build_all(force: bool = False) -> dict:
    res = {"oracle": build_oracle(force)}
    for name, fn in (("ref_cuda", build_ref_cuda), ("ref_cython", build_ref_cython)):
        try:
            res[name] = fn(force)
        except Exception as e:  # the reference side is optional; say why it is absent
            res[name] = None
            res[name + "_error"] = str(e)
    return res


if __name__ == "__main__":
    for k, v in build_all(force="--force" in sys.argv).items():
        print(k, "->", v)
