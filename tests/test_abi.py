"""The C-ABI library loads and exports every symbol include/scda_b200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "scda_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:int|void|size_t|unsigned long long)\s+(\w+)\s*\(", text, flags=re.M)
    assert len(names) >= 15
    return names


def test_library_is_built_and_loads():
    from scda_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run `python -m scda_b200.build`"
    lib = _lib.load()
    assert lib.scda_abi_version() == 2


def test_every_declared_symbol_is_exported_and_bound():
    from scda_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in scda_b200.h but not exported" % n
    # the Python binding covers exactly the declared set
    assert sorted(names) == sorted(_lib.SIGNATURES.keys())


def test_no_torch_types_in_header():
    text = open(os.path.join(ROOT, "include", "scda_b200.h")).read()
    for bad in ("at::", "torch::", "Tensor", "THC"):
        code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        assert bad not in code


def test_missing_library_fails_loudly(tmp_path):
    from scda_b200 import _lib
    import pytest
    with pytest.raises(_lib.ScdaLibraryError):
        _lib.load(str(tmp_path / "nope.so"))


def test_cpu_tensor_is_rejected_not_served():
    """No CPU path: a CPU tensor raises instead of silently computing somewhere else."""
    import pytest
    import torch
    from scda_b200 import _lib
    from scda_b200.extensions import RoIPool
    from scda_b200.extensions._roi_align.modules.roi_align import RoIAlign
    feat = torch.zeros(1, 4, 8, 8)
    rois = torch.tensor([[0., 0., 0., 31., 31.]])
    with pytest.raises(_lib.ScdaLibraryError):
        RoIPool(7, 7, 1 / 16.)(feat, rois)
    with pytest.raises(NotImplementedError):   # same as the reference, roi_align.py:30-31
        RoIAlign(7, 7, 1 / 16.)(feat, rois)


def test_product_does_not_import_oracle():
    """Nothing under scda_b200/ may reference the oracle (it is the checker)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "scda_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "liboracle" not in src and "oracle/" not in src, f


def test_header_is_plain_c(tmp_path):
    """include/scda_b200.h compiles on its own as C99 and as C++ (no CUDA, torch or STL header needed):
    it is the whole drop-in boundary a maintainer binds from cffi / ctypes / cgo."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest
        pytest.skip("no gcc")
    hdr = os.path.join(ROOT, "include", "scda_b200.h")
    src = tmp_path / "t.c"
    src.write_text('#include "%s"\nint main(void) { return scda_abi_version() == 0; }\n' % hdr)
    for lang, std in (("c", "-std=c99"), ("c++", "-std=c++11")):
        r = subprocess.run([gcc, "-x", lang, std, "-Wall", "-Werror", "-pedantic", "-fsyntax-only", str(src)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
