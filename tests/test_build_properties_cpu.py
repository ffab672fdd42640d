"""Static properties of the shipped library that the design depends on, read with cuobjdump (no GPU):

  * every instantiation of the tcgen05 contraction is built for at most 128 registers per thread -- at 320 threads
    per CTA that leaves registers on every SM sub-partition, which is what lets the update / bias-gradient kernels
    become resident beside a contraction CTA;
  * each of them really is a tcgen05 / TMEM / TMA kernel (UTCHMMA, LDTM, UTMALDG in its SASS), the lean variants
    carry the TMA reduce-add store (UTMAREDG) and the programmatic-dependent-launch wait (ACQBULK);
  * the NVLS update kernel uses multimem.ld_reduce (LDGMC ... ADD);
  * the code is sm_100a only (no PTX for a JIT to pick up on another architecture, no second arch)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "april_ann_b200", "libb200ann.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not (os.path.exists(CUOBJDUMP) and os.path.exists(LIB)),
                                reason="cuobjdump or the built library not available")


def run(*args):
    return subprocess.run([CUOBJDUMP, *args, LIB], capture_output=True, text=True).stdout


@pytest.fixture(scope="module")
def resources():
    out, name = {}, None
    for line in run("--dump-resource-usage").splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and name:
            out[name] = tuple(int(v) for v in m.groups())
    return out


@pytest.fixture(scope="module")
def sass():
    per, name = {}, None
    for line in run("-sass").splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            per[name] = set()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and name:
            per[name].add(m.group(1))
    return per


def test_contraction_variants_stay_at_128_registers(resources):
    tc = {k: v for k, v in resources.items() if "gemm_tc_kernel" in k}
    assert len(tc) == 20, sorted(tc)            # 4 layouts x (lean, lean+stamps, generic, generic+stamps) + 4 fused epilogues
    assert all(v[0] <= 128 for v in tc.values()), {k: v for k, v in tc.items() if v[0] > 128}


def test_contraction_is_tcgen05_tma(sass):
    tc = {k: v for k, v in sass.items() if "gemm_tc_kernel" in k}
    assert len(tc) == 20
    for k, ops in tc.items():
        for op in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "SYNCS", "ACQBULK"):
            assert op in ops, (k, op)
    # template arguments <A_KMAJOR, B_KMAJOR, ACT, DACT, LEAN, STAMPS>: ...Lb1E (LEAN) before the STAMPS flag
    lean = [k for k in tc if re.search(r"Lb1ELb[01]EEE", k)]
    assert len(lean) == 12
    assert all("UTMAREDG" in tc[k] for k in lean)


def test_multicast_update_kernel_uses_multimem(sass):
    mc = [k for k in sass if "dp_mc_update_kernel" in k]
    assert mc and all("LDGMC" in sass[k] for k in mc)


def test_only_sm_100a_code(resources):
    elf = run("-lelf")
    archs = set(re.findall(r"sm_(\d+a?)", elf))
    assert archs == {"100a"}, archs
    ptx = subprocess.run([CUOBJDUMP, "-lptx", LIB], capture_output=True, text=True)
    assert "PTX file" not in ptx.stdout or "No PTX file" in (ptx.stdout + ptx.stderr), ptx.stdout[:300]
