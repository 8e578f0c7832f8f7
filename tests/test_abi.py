"""CPU: the C-ABI shared library loads and exports every symbol include/mhaq_fq.h
declares; argument validation works without a GPU (no kernel is launched)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mhaq_fq.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mhaq_fq_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported():
    from mhaq_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 12
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in include/mhaq_fq.h but not exported"
    assert set(syms) == set(_lib.EXPORTED_SYMBOLS), "ctypes binding out of sync with the header"


def test_abi_version_and_info():
    from mhaq_b200 import _lib
    assert _lib.lib.mhaq_fq_abi_version() == 7
    assert b"sm_100a" in _lib.lib.mhaq_fq_build_info()


def test_geometry_is_a_pure_function_of_shape():
    from mhaq_b200 import _lib
    L = _lib.lib
    assert L.mhaq_fq_num_tasks(1, 4096) == 1
    assert L.mhaq_fq_num_tasks(1, 4097) == 2
    assert L.mhaq_fq_num_tasks(512, 4608) == 1024
    assert L.mhaq_fq_num_tasks(64, 576) == 64
    assert L.mhaq_fq_num_tasks(0, 10) == 0
    # streaming kernels: one task per 4096-element sub-tile
    assert L.mhaq_fq_num_tasks(1, 1 << 30) == 1 << 18
    # one finalize ticket per channel + the flat backward's record region (4 + 2048 x 12 words)
    flat = 4 + 2048 * 12
    assert L.mhaq_fq_ticket_count(1, 1 << 30, 1) == 1 + flat
    assert L.mhaq_fq_ticket_count(512, 4608, 512) == 512 + flat
    assert L.mhaq_fq_workspace_bytes(1, 1 << 30) >= (1 << 18) * 8 * 8


def test_argument_errors_without_gpu():
    from mhaq_b200 import _lib
    L = _lib.lib
    # null x / scale -> MHAQ_FQ_ENULL before anything touches CUDA
    assert L.mhaq_fq_fwd_f32(None, None, None, None, None, None, None, 0, 0, 0, 0, 0, 1, 8, 1, None, None) == -2
    dummy = ctypes.c_void_p(16)
    # bad stride
    assert L.mhaq_fq_fwd_f32(dummy, dummy, None, dummy, dummy, None, None, 2, 0, 0, 0, 0, 1, 8, 1, None, None) == -1
    # rows not divisible by channels
    assert L.mhaq_fq_fwd_f32(dummy, dummy, None, dummy, dummy, None, None, 1, 1, 0, 0, 0, 5, 8, 2, None, None) == -1
    # unknown parameter mode; ACT_LOG without log_act_q; ACT_LOG with channels
    assert L.mhaq_fq_fwd_f32(dummy, dummy, None, dummy, dummy, None, None, 0, 0, 0, 0, 7, 1, 8, 1, None, None) == -1
    assert L.mhaq_fq_fwd_f32(dummy, dummy, None, dummy, dummy, None, None, 0, 0, 0, 0, 1, 1, 8, 1, None, None) == -2
    assert L.mhaq_fq_fwd_f32(dummy, dummy, None, dummy, dummy, dummy, None, 0, 0, 0, 0, 1, 2, 8, 2, None, None) == -1
    # bad method
    assert L.mhaq_fq_bwd_f32(dummy, dummy, dummy, dummy, dummy, None, None, 0, 0, 0, 0, 0, 1, 8, 1, 9, 0,
                             None, 0, 0, None, None, dummy, None) == -1
    # missing tickets buffer
    assert L.mhaq_fq_bwd_finalize_f32(dummy, None, None, None, None, None, 0, 0, 0, 0, 0, 1, 8, 1, None, None, None,
                                      None, None) == -2
    with pytest.raises(RuntimeError, match="EINVAL"):
        _lib.check(-1, "x")


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle."""
    pkg = os.path.join(ROOT, "mhaq_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_estimator_ids_agree_between_header_enum_and_ops():
    """QNMethod values (the reference's, gdnsq_utils.py:9-13) are the ABI's `method` ids."""
    import re
    from mhaq_b200 import ops
    from mhaq_b200.quantization.gdnsq.gdnsq_utils import QNMethod, QMode
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                            "include", "mhaq_fq.h")).read()
    ids = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define MHAQ_FQ_(STE|EWGS|AEWGS|LSQ) (\d+)", hdr)}
    assert ids == {"STE": 0, "EWGS": 1, "AEWGS": 2, "LSQ": 3}
    assert {m.name: m.value for m in QNMethod} == ids == ops.METHOD_IDS
    assert [m.name for m in QMode] == ["NOISE_VAL", "ROUND_VAL", "SOURCE_VAL", "FLOAT_TRAIN_VAL"]
    assert [m.value for m in QMode] == [1, 2, 3, 4]


def test_header_is_plain_c():
    import subprocess
    hdr = os.path.join(ROOT, "include", "mhaq_fq.h")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-x", "c", hdr],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_plain_c_client_links_and_runs(tmp_path):
    """A C99 program (no C++, no torch) compiled against include/mhaq_fq.h and linked with the
    shared library runs the GPU-free entry points (version, geometry, argument validation)."""
    import subprocess
    from mhaq_b200 import _lib
    exe = str(tmp_path / "abi_client")
    libdir = os.path.dirname(_lib.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "c", "abi_client.c"), "-o", exe,
                        "-L", libdir, "-lmhaq_fq", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi_client OK" in r.stdout


def _header_param_counts():
    src = open(os.path.join(ROOT, "include", "mhaq_fq.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(mhaq_fq_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    return out


def test_binding_arities_match_the_header():
    """Every ctypes signature in mhaq_b200/_lib.py — and the maintainer-side stub printed in
    INTEGRATION.md — has exactly as many arguments as the C declaration."""
    from mhaq_b200 import _lib
    counts = _header_param_counts()
    assert set(counts) == set(_lib.EXPORTED_SYMBOLS)
    for name, (_, argtypes) in _lib._SIGNATURES.items():
        assert len(argtypes) == counts[name], f"{name}: binding {len(argtypes)} vs header {counts[name]}"
    # the stub of INTEGRATION.md section C, executed against the built library
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(import ctypes, torch\n.*?)\ndef fake_quant_fwd", md, flags=re.S).group(1)
    block = block.replace('ctypes.CDLL("libmhaq_fq.so")', f'ctypes.CDLL({_lib.LIB_PATH!r})')
    ns = {}
    exec(block, ns)
    for fn in ("mhaq_fq_fwd_f32", "mhaq_fq_bwd_f32", "mhaq_fq_bwd_finalize_f32", "mhaq_fq_workspace_bytes"):
        assert len(getattr(ns["L"], fn).argtypes) == counts[fn], f"INTEGRATION.md stub: {fn}"
    assert (ns["LINEAR"], ns["ACT_LOG"], ns["WEIGHT_LOG"]) == (0, 1, 2)
