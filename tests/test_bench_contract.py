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


def test_clustalo_leg_runs_the_wrappers_argv_when_a_binary_is_on_path(tmp_path, monkeypatch):
    """with a `clustalo` on PATH the CPU-baseline leg runs the reference wrapper's argv
    (Core/ClustalO.cpp:51) + --full --distmat-out on configs[0] and reports wall time and the rank correlation of its
    distances with ours.  No such binary exists in the image, so a stand-in that honours exactly that argv (and
    writes a distance matrix in clustalo's format) exercises the leg end to end; without one it reports absence."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setenv("PATH", "/nonexistent")
    assert bench.clustalo_leg(["ACD", "ACE"], np.zeros(1)) == {"available": False, "note": "ClustalO not available in image"}
    fake = tmp_path / "clustalo"
    fake.write_text(f"""#!{sys.executable}
import sys
args = sys.argv[1:]
assert args[:4] == ["--force", "-v", "--outfmt=fa", "--output-order=tree-order"], args      # ClustalO.cpp:51
fin, fout = args[args.index("-i") + 1], args[args.index("-o") + 1]
assert "--full" in args
mat = [a.split("=", 1)[1] for a in args if a.startswith("--distmat-out=")][0]
names, seqs = [], []
for line in open(fin):
    line = line.strip()
    if line.startswith(">"):
        names.append(line[1:].split()[0]); seqs.append("")
    elif line:
        seqs[-1] += line
def d(a, b):                       # a crude composition distance: enough for a rank correlation
    return sum(abs(a.count(c) - b.count(c)) for c in set(a + b)) / max(len(a) + len(b), 1)
with open(mat, "w") as f:
    f.write(f"{{len(names)}}\\n")
    for i, n in enumerate(names):
        f.write(n + " " + " ".join(f"{{d(seqs[i], s):.6f}}" for s in seqs) + "\\n")
open(fout, "w").write("".join(f">{{n}}\\n{{s}}\\n" for n, s in zip(names, seqs)))
""")
    fake.chmod(0o755)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.path.dirname(sys.executable))
    from tweakseq_b200 import synth
    seqs = synth.protein(12, 40, 3, family=True)
    ours = np.array([sum(abs(a.count(c) - b.count(c)) for c in set(a + b)) / (len(a) + len(b))
                     for i, a in enumerate(seqs) for b in seqs[i + 1:]])
    leg = bench.clustalo_leg(seqs, ours)
    assert leg["available"] and leg["returncode"] == 0 and leg["n"] == 12 and leg["seconds"] > 0
    assert "--full" in leg["argv"] and "--distmat-out=" in leg["argv"] and "--output-order=tree-order" in leg["argv"]
    assert abs(leg["spearman_vs_ours"] - 1.0) < 1e-9, leg
