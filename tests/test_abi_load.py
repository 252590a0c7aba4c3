"""CPU: libb200rec.so loads and exports every symbol declared in include/b200rec.h; the binding table in
b200rec/_abi.py lists exactly those symbols; compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import REPO
from b200rec import _abi


def _declared():
    text = open(os.path.join(REPO, "include", "b200rec.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200rec_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _abi.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert sorted(_abi.PROTOTYPES) == names
    assert lib.b200rec_version() >= 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from b200rec import ops
    with pytest.raises(_abi.B200RecError):
        ops.score_dense(torch.zeros(4, 64), torch.zeros(2, dtype=torch.int64), torch.zeros(4, 64))
    import model
    import dataset
    from b200rec import synth
    ds = dataset.get_dataset({'name': 'SyntheticDataset', 'device': 'cpu', 'graph': synth.generate(50, 60, 400, seed=1)})
    with pytest.raises(_abi.B200RecError):
        model.get_model({'name': 'LightGCN', 'embedding_size': 64, 'n_layers': 2, 'device': 'cpu'}, ds)


def test_argument_errors_are_codes_with_messages():
    """every entry point validates before it launches: a bad call returns a negative code and leaves a message in
    b200rec_last_error() -- checked here without a GPU (validation comes before any CUDA call)"""
    lib = _abi.load()
    null = C.c_void_p(0)

    def err():
        return lib.b200rec_last_error().decode()

    rc = lib.b200rec_spmm_f32(None, null, 64, null, C.c_float(1.0), null, null, null, C.c_float(1.0), null)
    assert rc < 0 and "null operand" in err()
    csr = _abi.CsrStruct()
    x = np.zeros(64, dtype=np.float32)
    rc = lib.b200rec_spmm_f32(C.byref(csr), x.ctypes.data, 64, null, C.c_float(1.0), null, null, null, C.c_float(1.0), null)
    assert rc < 0 and "no output" in err()
    rc = lib.b200rec_spmm_f32(C.byref(csr), x.ctypes.data, 48, null, C.c_float(1.0), x.ctypes.data, null, null, C.c_float(1.0),
                              null)
    assert rc < 0 and "8/16/32/64/128/256" in err()
    rc = lib.b200rec_infonce_fwd_bwd(x.ctypes.data, x.ctypes.data, null, 1, 4, 16, C.c_float(0.0), C.c_float(1.0), x.ctypes.data,
                                     x.ctypes.data, x.ctypes.data, x.ctypes.data, null)
    assert rc < 0 and "temperature" in err()
    rc = lib.b200rec_infonce_fwd_bwd(x.ctypes.data, x.ctypes.data, null, 1, 4, 24, C.c_float(0.1), C.c_float(1.0), x.ctypes.data,
                                     x.ctypes.data, x.ctypes.data, x.ctypes.data, null)
    assert rc < 0 and "embedding size" in err()
    step = np.zeros(1, dtype=np.int64)
    rc = lib.b200rec_adam_step(x.ctypes.data + 4, x.ctypes.data, x.ctypes.data, x.ctypes.data, 8, 1e-3, 0.9, 0.999, 1e-8,
                               step.ctypes.data, null)
    assert rc < 0 and "alignment" in err()
    assert lib.b200rec_infonce_workspace_floats(0, 64) == 0 and lib.b200rec_infonce_workspace_floats(2048, 64) > 2 * 2048 * 64


def test_library_is_sm100a_and_carries_tensor_core_code():
    """the built library holds sm_100a cubins only, and the evaluation kernel really is tcgen05 / TMA / TMEM code
    (SASS: UTCHMMA = tcgen05.mma, UTMALDG = TMA tile load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit)"""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    elf = subprocess.run([cuobjdump, "-lelf", _abi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    cubins = [line.split()[-1] for line in elf.splitlines() if "cubin" in line]
    assert cubins and all(".sm_100a." in c for c in cubins), cubins
    sass = subprocess.run([cuobjdump, "-sass", _abi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR"):
        assert mnemonic in sass, mnemonic


def test_csr_struct_mirror_matches_header_and_integration_doc():
    """b200rec_csr: field names, order and C types in include/b200rec.h == the ctypes mirror in b200rec/_abi.py == the stub
    shown in INTEGRATION.md (a reordered field would silently shift every pointer)"""
    text = open(os.path.join(REPO, "include", "b200rec.h")).read()
    body = text[text.index("typedef struct {"):text.index("} b200rec_csr;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"^\s*(const\s+)?(int32_t|float|void)\s*(\*?)\s*([a-z0-9_]+);", body, flags=re.M)
    header = [(name, "ptr" if star else ctype) for _, ctype, star, name in fields]
    mirror = [(n, "ptr" if t is C.c_void_p else {C.c_int32: "int32_t"}[t]) for n, t in _abi.CsrStruct._fields_]
    assert header == mirror and len(header) >= 20
    doc = open(os.path.join(REPO, "INTEGRATION.md")).read()
    stub = doc[doc.index("class b200rec_csr(ctypes.Structure)"):doc.index("def get_rep(self):")]
    doc_fields = re.findall(r'\("([a-z0-9_]+)", ctypes\.(c_int32|c_void_p)\)', stub)
    assert [(n, "ptr" if t == "c_void_p" else "int32_t") for n, t in doc_fields] == mirror
