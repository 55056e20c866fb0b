"""CPU-only checks of the drop-in boundary: the library builds, loads, exports every symbol the header declares, and
fails loudly (no fallback) when there is no GPU or an argument is wrong."""
import ctypes
import os
import re

import pytest
import torch

from ml_inference_optimizer_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported(built_lib):
    header = open(os.path.join(ROOT, "include", "b200_attn_mlp.h")).read()
    declared = sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(built_lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == declared


def test_header_cites_reference_interfaces():
    header = open(os.path.join(ROOT, "include", "b200_attn_mlp.h")).read()
    for cite in ("flash_attention_kernels.py:1150", "attention_kernels.py:1206", "attention_kernels.py:1314",
                 "mlp_kernels.py:648", "tensor_parallel.py:173", "tensor_parallel.py:296-308"):
        assert cite in header


def test_integration_notes_place_every_symbol():
    header = open(os.path.join(ROOT, "include", "b200_attn_mlp.h")).read()
    notes = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [s for s in sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", header))) if s not in notes]
    assert not missing, f"INTEGRATION.md does not say where these entry points bind: {missing}"


def test_version_and_pure_host_queries(built_lib):
    assert b"sm_100a" in built_lib.b200_version()
    # wide problem: no split-K; the intermediate + the single-launch kernel's counters (2 per 256-row block + 1, 256-aligned)
    assert built_lib.b200_fused_mlp_workspace_bytes(32768, 4096, 11008) == 32768 * 11008 * 2 + 1280
    assert built_lib.b200_linear_act_workspace_bytes(32768, 4096, 11008, 4) == 0
    assert built_lib.b200_fa_decode_workspace_bytes(4, 32, 8, 128, 8192, 1) == 0
    assert built_lib.b200_fa_decode_workspace_bytes(4, 32, 8, 128, 8192, 4) == 4 * 32 * 4 * 129 * 4


def test_invalid_arguments_return_error_codes(built_lib):
    s3 = _lib.strides3((1024, 128, 128))
    rc = built_lib.b200_fa_fwd(None, None, None, None, None, 1, 128, 128, 1, 1, 128, s3, s3, s3, s3, 0.1, 0, 0, None, 0, None)
    assert rc == -1 and "NULL" in _lib.last_error()
    dummy = ctypes.c_void_p(256)
    rc = built_lib.b200_fa_fwd(dummy, dummy, dummy, dummy, None, 1, 128, 128, 3, 2, 128, s3, s3, s3, s3, 0.1, 0, 0, None, 0, None)
    assert rc == -1 and "multiple of Hkv" in _lib.last_error()
    rc = built_lib.b200_fa_fwd(dummy, dummy, dummy, dummy, None, 1, 128, 128, 2, 2, 100, s3, s3, s3, s3, 0.1, 0, 0, None, 0, None)
    assert rc == -1 and "head_dim" in _lib.last_error()
    rc = built_lib.b200_linear_act(dummy, 60, dummy, None, None, None, dummy, 64, 8, 60, 64, 0, None, 0, 0, None)
    assert rc == -1 and "multiples of 8" in _lib.last_error()
    rc = built_lib.b200_fused_mlp(dummy, 64, dummy, None, None, None, dummy, None, dummy, 64, 8, 64, 128, 64, 1, None, 0, 0, None)
    assert rc == -5 and "workspace" in _lib.last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from ml_inference_optimizer_b200 import ops

    assert _lib.load().b200_arch_ok() < 0  # B200_ERR_NO_DEVICE
    q = torch.randn(1, 128, 1, 128, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.flash_attn_fwd(q, q, q)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.fused_mlp(torch.randn(4, 64, dtype=torch.bfloat16), torch.randn(128, 64, dtype=torch.bfloat16), None,
                      torch.randn(64, 128, dtype=torch.bfloat16), None, "relu")


def test_product_package_never_imports_the_oracle():
    import re as _re

    for top in ("ml_inference_optimizer_b200", "kernels", "parallelism", "baseline", "ml_inference_optimizer"):
        pkg = os.path.join(ROOT, top)
        for dirpath, _, files in os.walk(pkg):
            for f in files:
                if f.endswith(".py"):
                    text = open(os.path.join(dirpath, f)).read()
                    assert not _re.search(r"^\s*(from|import)\s+oracle\b", text, _re.M), f"{top}/{f} imports the oracle"


def test_build_is_keyed_on_source_content(built_lib, monkeypatch, tmp_path):
    """The library on disk must have been built from exactly the current sources: the stamp next to it is a hash of the
    sources, headers and flags; a stale or missing stamp (a prebuilt .so travelling with a snapshot, a checkout that reset
    mtimes) forces a rebuild instead of being trusted."""
    from ml_inference_optimizer_b200 import build as b

    assert b.LIB_PATH.exists() and b.STAMP_PATH.exists()
    assert b.STAMP_PATH.read_text().strip() == b.source_stamp()
    assert not b.needs_build()
    stale = tmp_path / "source_stamp.txt"
    stale.write_text("0" * 64 + "\n")
    monkeypatch.setattr(b, "STAMP_PATH", stale)
    assert b.needs_build()
    monkeypatch.setattr(b, "STAMP_PATH", tmp_path / "missing.txt")
    assert b.needs_build()
    monkeypatch.setenv("B200_EXTRA_NVCC_FLAGS", "-DSOMETHING_ELSE")   # other flags = another binary
    assert b.source_stamp() != (b.PKG_DIR / "build" / "source_stamp.txt").read_text().strip()


def test_lib_path_override_is_explicit(monkeypatch, tmp_path):
    """B200_LIB_PATH (developer A/B runs) must name an existing file: a wrong path raises instead of silently loading the
    default library."""
    import importlib

    from ml_inference_optimizer_b200 import _lib as L

    monkeypatch.setenv("B200_LIB_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setattr(L, "_lib", None)
    with pytest.raises(RuntimeError, match="missing"):
        L.load()
    monkeypatch.delenv("B200_LIB_PATH")
    monkeypatch.setattr(L, "_lib", None)
    assert L.load() is not None


def test_co_residency_budget_of_the_tensor_parallel_overlap(built_lib):
    """The in-switch all-reduce CTAs (K6) are meant to run UNDERNEATH the persistent GEMM CTAs of the next token chunk: on
    one SM, 256 GEMM threads + 256 all-reduce threads must fit the 64K registers (allocation granularity 8 per thread). A
    refactor that silently grows the GEMM kernel's registers (it happened: 209 -> 250 cost 0.45 ms of a 4.3 ms TP step) or
    makes it reserve all 16 named barriers breaks the overlap without failing any numerical test — so it is checked here."""
    import shutil
    import subprocess

    from ml_inference_optimizer_b200 import build as b

    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([tool, "-res-usage", str(b.LIB_PATH)], capture_output=True, text=True).stdout
    regs = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and name:
            regs[name] = int(m.group(1))
    up8 = lambda r: (r + 7) // 8 * 8
    gemm = {k: v for k, v in regs.items() if "gemm_act_pair_kernelILi4E" in k or "gemm_act_pair_kernelILi0E" in k}
    gemm = {k: v for k, v in gemm.items() if "Li1EEE" in k}          # the one-epilogue-warpgroup builds the TP path launches
    ar = {k: v for k, v in regs.items() if "tp_allreduce_kernel" in k}
    assert gemm and ar, "kernels not found in the library"
    worst = max(up8(g) for g in gemm.values()) * 256 + max(up8(a) for a in ar.values()) * 256
    assert worst <= 65536, (gemm, ar)


def test_measured_heuristics_are_pinned(built_lib):
    """The split rules were fitted to sweeps on the GPU (profiles/r2_decode_splits.jsonl, profiles/r2_gemm_splits.jsonl);
    these are the choices the sweeps found best, pinned so that a later edit of the rules has to look at the data again.
    (Pure host functions: 148 SMs are assumed when no device is present.)"""
    ns = built_lib.b200_fa_decode_num_splits
    #        B  Hq Hkv   D  context                                  measured optimum
    assert ns(32, 16, 1, 128, 8192) == 8      # MQA, 32 CTAs per split: one wave of 256 CTAs (46 us vs 79 us at 32 splits)
    assert ns(8, 32, 8, 128, 8192) == 4       # 64 x 4 = 256 CTAs (64 us; 5 splits = 1.08 waves: 89 us)
    assert ns(1, 32, 8, 128, 32768) == 32
    assert ns(64, 32, 8, 128, 8192) == 4      # C4: 512 CTAs = 1.73 waves -> 4 splits fill the last wave
    assert ns(64, 32, 32, 128, 8192) == 1     # C3 MHA: 2048 CTAs, never split
    assert ns(64, 16, 4, 64, 8192) == 8       # D = 64 GQA (3 CTAs per SM): under-filled single wave, bandwidth-sized -> several waves
    assert ns(1, 12, 12, 64, 64) == 1         # GPT-2 decode at short context: >= 256 keys per split
    ws = built_lib.b200_linear_act_workspace_bytes
    assert ws(64, 4096, 11008, 4) == 0                                  # SwiGLU up+gate at T=64: 86 tiles, no split (37.9 vs 42.3 us)
    assert ws(64, 11008, 4096, 0) == 9 * 64 * 16 * 256 * 4              # down projection at T=64: 16 tiles x 9 splits
    assert ws(8, 4096, 11008, 4) == 5 * 8 * 86 * 256 * 4                # T=8: small partials -> 5 splits (37.4 vs 44.8 us)
    assert ws(32768, 4096, 11008, 4) == 0                               # prefill: never split
