"""CPU suite: the C-ABI library loads and exports every symbol include/dfrl.h declares, the
host-side logic of bench.py, and the multi-rank plumbing under gloo (world_size 2)."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import refcases

ROOT = refcases.ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "dfrl.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dfrl_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import dependence_free_rl_b200 as D
    names = _declared_symbols()
    assert len(names) > 60
    for n in names:
        assert hasattr(D._lib.lib, n), f"libdfrl_b200.so does not export {n}"
        assert n in D._lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert set(D._lib.PROTOTYPES) == set(names)
    assert b"sm_100a" in D._lib.lib.dfrl_version()


def test_header_cites_reference_interfaces():
    src = open(os.path.join(ROOT, "include", "dfrl.h")).read()
    for cite in ["bin_packing.h:53-64", "nn.h:72-79", "rl.h:45-74", "policy_gradient.h:125-147",
                 "nn.h:616-698", "tensor.cc:467-470"]:
        assert cite in src


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import dependence_free_rl_b200 as D
    with pytest.raises(D._lib.DfrlError, match="no CPU fallback"):
        D.Context(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dependence_free_rl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.replace("test oracle", ""), f"{f} mentions the oracle"


def test_bench_flop_model_matches_survey():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.flops_per_env_step() == 290592
    assert bench.flops_per_env_step([128, 256, 256, 256, 32], [128, 256, 256, 256, 1]) == 7556224


def test_gloo_world_size_2_plumbing():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29561")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29561",
           os.path.join(ROOT, "tests", "gloo_worker.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    assert json.loads(line) == {"ok": True, "world": 2}


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "1", "--ref-envs", "16"],
                         capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0


def _fused_object():
    return os.path.join(ROOT, "dependence_free_rl_b200", "csrc", ".out", "fused.o")


@pytest.mark.skipif(not os.path.exists(_fused_object()), reason="library built elsewhere (no object files)")
def test_fused_kernels_are_tcgen05_code():
    """SASS of the fused kernels (cross-compiled here, no GPU needed): tensor-core MMAs with operands
    in shared AND tensor memory, TMEM loads / stores, mbarrier completion, the bulk copy of the
    parameter vector -- and no wgmma / mma.sync fallback."""
    sass = subprocess.run(["cuobjdump", "-sass", _fused_object()], capture_output=True, text=True, check=True).stdout
    for mnemonic in ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "SYNCS.PHASECHK", "UBLKCP"]:
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass and "HGMMA" not in sass
    # tcgen05.mma with the A operand in tensor memory ("tmem[" as a source operand of UTCHMMA)
    assert re.search(r"UTCHMMA\s+tmem\[", sass), "no tcgen05.mma with a TMEM A operand"


@pytest.mark.skipif(not os.path.exists(_fused_object().replace("fused.o", "fused.ptxas.log")),
                    reason="no ptxas log (library built elsewhere)")
def test_fused_kernels_do_not_spill():
    """The fused kernels run with the whole shared-memory carve-out, i.e. without an L1: a register
    spill costs an L2 round trip (the critic step lost 20 % to 136 spilled words). Guard: at most
    32 bytes of spill stores per kernel (the 576-thread critic step is capped at 96 registers; its small-batch
    variant with the in-kernel end pass spills 5 words)."""
    log = open(_fused_object().replace("fused.o", "fused.ptxas.log")).read()
    entries = re.findall(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores", log)
    fused = [(n, int(st)) for n, _, st in entries if "fused_" in n and "reduce" not in n]
    assert len(fused) >= 8
    for name, spill in fused:
        assert spill <= 32, (name, spill)
