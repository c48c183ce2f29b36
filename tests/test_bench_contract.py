"""bench.py contract pieces that can be checked without a GPU: deterministic synthetic inputs and the
`--impl reference` arm (the CPU oracle timed as the reference's stand-in)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_synthetic_genomes_are_deterministic_and_in_range():
    a = bench.synthetic_genomes(0, 8, 48)
    b = bench.synthetic_genomes(4, 4, 48)
    assert a.dtype == np.uint8 and a.max() <= 10 and np.array_equal(a[4:], b)      # genome g depends only on its global index
    assert len({bytes(r) for r in a}) == 8


def test_trial_list_and_attempt_count():
    tsz, tsd = bench.trial_list(bench.SIZES_C2, bench.SEEDS_C2)
    assert tsz[:5] == [(6, 4)] * 5 and tsd[:6] == [1, 2, 3, 4, 5, 1] and len(tsz) == 15
    assert bench.attempts_per_generation() == 50 * 5 * (61 ** 3 + 169 ** 3 + 271 ** 3)


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-pop", "1"], capture_output=True, text=True, check=True, cwd=ROOT).stdout
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "particle-move attempts/s" and line["unit"] == "attempts/s"
    assert line["higher_is_better"] is True and line["value"] > 1e6
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("C2")
