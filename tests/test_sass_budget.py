"""Static checks of the built library's device code (cuobjdump, no GPU needed): the hot kernels are sm_100a DPX / TMA
code, stay inside their register budgets without spilling, and their inner loops keep the instruction count per packed
cell that DESIGN.md quotes.  A regression here (a spill in the row loop, a lost DPX form, a variant that no longer
fits three CTAs per SM) would otherwise only show up as a slower bench line on the GPU box."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tweakseq_b200", "libtsqb200.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(SO),
                                reason="needs cuobjdump and the built library")


@pytest.fixture(scope="module")
def usage():
    """kernel (demangled-ish mangled name) -> (registers, stack bytes, shared bytes)"""
    txt = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True, check=True).stdout
    out, name = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and name:
            out[name] = tuple(int(x) for x in m.groups())
            name = None
    return out


def pick(usage, *parts):
    hits = [k for k in usage if all(p in k for p in parts)]
    assert len(hits) == 1, (parts, hits)
    return usage[hits[0]]


def test_only_sm_100a_code_is_in_the_library():
    txt = subprocess.run(["cuobjdump", "-lelf", SO], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", txt))
    assert archs == {"100a"}, archs


def test_register_budgets_and_no_spills_in_the_hot_kernels(usage):
    # packed inter-task kernel: three CTAs of 128 threads per SM need <= 168 registers; the default strip widths of
    # configs[1] / configs[2] (44, 50) and everything below must not spill into the row loop's way
    for K, bounds in ((30, 4), (32, 4), (36, 3), (40, 3), (44, 3), (50, 3), (56, 3)):
        for nge in ("Lj0E", "Lj65537E"):
            reg, stack, _ = pick(usage, f"gotoh16_kernelILi{K}ELi128ELi{bounds}E", nge)
            assert reg <= (128 if bounds == 4 else 168), (K, reg)
            if K <= 44:
                assert stack == 0, (K, nge, stack)
            else:
                assert stack <= 64, (K, nge, stack)        # task-level values only; none in the row loop
    # packed wavefront kernel, nucleotide column block (24 columns per lane, 3 CTAs per SM): no stack at all
    for nge in ("Lj0E", "Lj65537E"):
        reg, stack, _ = pick(usage, "wave16_kernelILi24ELi128ELi3E", nge)
        assert reg <= 168 and stack == 0, (reg, stack)
    # progressive-alignment merge: one variant, 128 registers, the tile's scores and edges in registers
    reg, stack, _ = pick(usage, "msa_merge_kernelILi512E")
    assert reg <= 128 and stack == 0, (reg, stack)


def sass_loop(name, cells, block=False):
    cmd = [sys.executable, os.path.join(ROOT, "tools", "sass_inner_loop.py"), name, str(cells)] + (["block"] if block else [])
    txt = subprocess.run(cmd, capture_output=True, text=True, check=True, cwd=ROOT).stdout
    per = float(re.search(r"-> ([0-9.]+) instructions per packed cell", txt).group(1))
    hist = {m.group(2): int(m.group(1)) for m in re.finditer(r"^\s+(\d+)\s+(\S+)$", txt, flags=re.M)}
    return per, hist


def test_inner_loops_keep_their_instruction_budget():
    # gotoh16, K = 44: a row pair is 88 packed cells: 3 DPX + 2 adds + 1 LDS each and ~26 around them
    per, hist = sass_loop("gotoh16_kernelILi44ELi128ELi3ELj65537E", 88)
    assert per <= 6.45, per
    assert hist.get("VIMNMX3.U16x2") == 88 and hist.get("VIADDMNMX.U16x2") == 176, hist
    # wave16, 24 columns per lane: the four-row body is 96 packed cells with one LDS.128 per four
    per, hist = sass_loop("wave16_kernelILi24ELi128ELi3ELj65537E", 96, block=True)
    assert per <= 5.6, per
    assert hist.get("VIMNMX3.U16x2") == 96 and hist.get("VIADDMNMX.U16x2") == 192 and hist.get("LDS.128") == 24, hist


def test_tma_and_mbarrier_forms_are_present():
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN3tsq13wave16_kernelILi24ELi128ELi3ELj65537EEEvNS_9W16ParamsE", SO],
                         capture_output=True, text=True).stdout
    assert "UBLKCP.S.G" in txt and "SYNCS.ARRIVE.TRANS64" in txt and "SYNCS.PHASECHK.TRANS64.TRYWAIT" in txt
    assert "WARPSYNC.COLLECTIVE" not in txt.split("SHFL.UP")[0][-2000:]      # the shuffles are not behind a collective path
