"""bench.py's reference arm (the CPU oracle on all host cores) runs without a GPU: check that it
prints one JSON line with the contract's keys.  The GPU arm is exercised by the driver."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GCUPS" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "configs[1]" in line["config"]["workload"]
    assert line["clustalo"]            # either a path or the explicit "not available" note


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_parity_block_checks_the_delivered_result_against_the_oracle():
    """bench.py's in-line parity object (sampled pairs + first and last rows), on the CPU: clean on the oracle's
    own matrix, and it sees a single wrong score or distance wherever it sits in the checked set."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from oracle import pyoracle as o
    from tweakseq_b200 import synth
    seqs = synth.protein(40, 60, 9)
    enc = [o.encode(s) for s in seqs]
    mat = o.matrix(0)
    s, _ = o.all_pairs(enc, mat, 11, 1, nthreads=2)
    d = o.distances(s, np.array([o.self_score(e, mat) for e in enc], dtype=np.int32))
    n = len(seqs)
    blk = bench.parity_block(seqs, 0, s, d, 10_000, 1, 2)
    assert blk["mismatches"] == 0 and blk["distance_mismatches"] == 0
    assert blk["pairs_checked"] == n * (n - 1) // 2 and blk["first_row_pairs"] == n - 1    # sample >= population: everything
    blk = bench.parity_block(seqs, 0, s, None, 50, 2, 2)
    assert blk["mismatches"] == 0 and "distance_mismatches" not in blk and n - 1 + 6 <= blk["pairs_checked"] <= 70 + n - 1 + 6
    for bad in (0, n - 2, len(s) - 1):           # first pair, end of the first row, the very last pair: always checked
        s2, d2 = s.copy(), d.copy()
        s2[bad] += 1
        d2[bad] = np.nextafter(d2[bad], 2.0)
        blk = bench.parity_block(seqs, 0, s2, d2, 10, 3, 2)
        assert blk["mismatches"] == 1 and blk["distance_mismatches"] == 1
