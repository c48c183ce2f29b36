"""Host-side mirror of the reference's GACore + CLI (evosops_b200/host): GA operators, trial lists, genome
file parsing, Rust-compatible log formatting and the genomic_data_*.log format.  CPU tests use the stub
evaluator binary (evosops_hosttest); the GPU test runs the real `evosops` CLI end to end."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "evosops_b200", "host")


@pytest.fixture(scope="module")
def built():
    from evosops_b200 import _build
    _build.build()
    subprocess.check_call(["make", "-C", HOST, "-s"])
    return HOST


def run(built, *args, **kw):
    return subprocess.run([os.path.join(built, "evosops_hosttest")] + [str(a) for a in args], capture_output=True, text=True,
                          check=True, **kw).stdout


def test_rust_float_formatting(built):
    # println!("{}", 0.65f32) / 0.08f64 / {:?} of 1.0 / {} of 1.0 / shortest round-trip f64
    assert run(built, "fmt").split() == ["0.65", "0.08", "1.0", "1", "0.5194300637633971"]


def test_genome_file_parsing_pops_last_char(built, tmp_path, fixtures):
    # main.rs:199-210: the parser strips brackets/spaces and pop()s the LAST character unconditionally,
    # so a file must end with exactly one extra character (normally the newline)
    g = np.asarray(fixtures["agg_genome_evolved"], dtype=np.uint8)
    text = str(g.tolist())
    p = tmp_path / "g.txt"
    p.write_text(text + "\n")
    assert [int(v) for v in run(built, "parse", p).split()] == g.reshape(-1).tolist()
    p.write_text(text)                     # no trailing char: the final ']' is stripped first, then the last digit is popped
    got = [int(v) for v in run(built, "parse", p).split()]
    assert len(got) == 47 and got == g.reshape(-1).tolist()[:47]      # the last gene is lost, exactly as upstream


def test_trial_list_orders(built):
    # base_ga.rs:356-362 sizes-major / seeds-minor; loco_ga.rs:384-389 zip(sizes, seeds) each repeated len(seeds) times
    agg = run(built, "trials", "agg").split()
    assert agg == ["(6,3,1):11", "(6,3,1):22", "(6,3,1):33", "(10,7,2):11", "(10,7,2):22", "(10,7,2):33",
                   "(13,9,3):11", "(13,9,3):22", "(13,9,3):33"]
    loco = run(built, "trials", "loco").split()
    assert loco == ["(6,3,1):11"] * 3 + ["(10,7,2):22"] * 3 + ["(13,9,3):33"] * 3


@pytest.mark.parametrize("beh,L", [("agg", 48), ("sep", 600), ("loco", 144)])
def test_ga_run_log_and_genomic_data(built, tmp_path, beh, L):
    pop, gens = 12, 8
    out = run(built, "ga", beh, pop, gens, 10, 0.08, 7, tmp_path)
    assert out.count("Starting Gen:") == gens and out.count("Generation Elapsed Time: ") == gens
    assert out.count("Avg. Fitness -> ") == gens and out.count("Population diversity -> ") == gens
    name = {"agg": "Genome", "sep": "SepGenome", "loco": "LocoGenome"}[beh]
    best = re.findall(r"Best Genome -> %s \{ string: (\[.*\]), fitness: (\d\.\d{5}) \}" % name, out)
    assert len(best) == gens
    assert np.asarray(eval(best[0][0])).size == L            # Debug-printed nested array parses back
    if beh == "agg":
        assert "Mutation Rate:" not in out                    # base_ga.rs prints no adaptive state
    else:
        assert out.count("Mutation Rate:") == gens and out.count("Diversity State:") == gens
        assert "Thresholds: UPPER: " in out and "Weights: 0.65 x Total Edges" in out
    # genomic_data_<tag>_<rand>.log: per generation, per genome: L gene bytes + f64 fitness BIG-endian
    tag = {"agg": "Agg", "sep": "Sep", "loco": "Loco"}[beh]
    data = open(os.path.join(str(tmp_path), "genomic_data_%s_424242.log" % tag), "rb").read()
    assert len(data) == gens * pop * (L + 8)
    rec = data[:L + 8]
    fit = struct.unpack(">d", rec[L:])[0]
    assert 0.0 < fit < 1.0 and max(rec[:L]) <= 10
    avg0 = float(re.search(r"Avg. Fitness -> ([0-9.]+)", out).group(1))
    fits = [struct.unpack(">d", data[k * (L + 8) + L:(k + 1) * (L + 8)])[0] for k in range(pop)]
    assert abs(np.mean(fits) - avg0) < 1e-12                  # the file holds the evaluated population (written before selection)


def test_ga_selection_improves_stub_fitness(built, tmp_path):
    out = run(built, "ga", "agg", 40, 40, 10, 0.08, 3, tmp_path)
    avgs = [float(v) for v in re.findall(r"Avg. Fitness -> ([0-9.]+)", out)]
    assert avgs[-1] > avgs[0] + 0.1


def test_agg_cache_skips_known_genomes_sep_does_not(built, tmp_path):
    # agg consults genome_cache (base_ga.rs:366-398); sep has it commented out (sep_ga.rs:391-417)
    T = 6
    agg = run(built, "ga", "agg", 30, 25, 2, 0.0, 11, tmp_path)          # gran 2 + no mutation -> many duplicates
    sep = run(built, "ga", "sep", 30, 25, 2, 0.0, 11, tmp_path)
    n_agg = int(re.search(r"INSTANCES (\d+)", agg).group(1))
    n_sep = int(re.search(r"INSTANCES (\d+)", sep).group(1))
    assert n_sep == 30 * 25 * T
    assert n_agg <= 30 * 25 * T


def test_adaptive_mutation_rate_switches_on_low_diversity(built, tmp_path):
    # sep_ga.rs:538-551: INIT -> LOWER_HIT multiplies the rate by 10 when the 10-generation mean diversity <= 0.08
    out = run(built, "ga", "sep", 20, 40, 10, 0.0005, 5, tmp_path)
    rates = [float(v) for v in re.findall(r"Mutation Rate:([0-9.e-]+)", out)]
    states = re.findall(r"Diversity State:(\w+)", out)
    assert states[0] == "INIT" and "LOWER_HIT" in states
    k = states.index("LOWER_HIT")
    assert abs(rates[k] - 10 * rates[k - 1]) < 1e-12


def test_cli_argument_errors(built):
    exe = os.path.join(built, "evosops")
    r = subprocess.run([exe, "-b", "agg"], capture_output=True, text=True)
    assert r.returncode == 2 and "--exp is required" in r.stderr
    r = subprocess.run([exe, "-b", "xyz", "-e", "ga", "-k", "(6,3)", "-s", "1"], capture_output=True, text=True)
    assert r.returncode == 2 and "invalid value" in r.stderr


@pytest.mark.gpu
def test_cli_end_to_end_on_gpu(built, tmp_path, fixtures, genomes):
    from oracle import oracle as O
    exe = os.path.join(built, "evosops")
    gfile = tmp_path / "genome.txt"
    gfile.write_text(str(np.asarray(fixtures["agg_genome_evolved"]).tolist()) + "\n")
    out_dir = tmp_path / "output"
    r = subprocess.run([exe, "-b", "agg", "-e", "gm", "-k(13,9)", "-s", "62813", "--gran", "10", "--path", str(gfile),
                        "--out", str(out_dir), "--ga-seed", "9"], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    m = re.search(r'Please check: "(Agg_GM_1_sizes_1_trials_gran_10_\d+)" file', r.stdout)
    log = open(os.path.join(str(out_dir), m.group(1) + ".log")).read()
    assert "Target Behavior: Agg" in log and "Experiment Type: GM" in log and 'Particle Sizes: ["(13,9)"]' in log
    grids = re.findall(r"SOPS grid\n((?: [0-9 ]+\n){27})", log)
    assert len(grids) == 4                                    # initial, after step n, after step n^2, final
    first = np.asarray([[int(v) for v in row.split()] for row in grids[0].strip("\n").split("\n")], dtype=np.uint8)
    assert np.array_equal(first, np.asarray(fixtures["agg_13_9_snapshots"][0], dtype=np.uint8))   # the reference's own step-0 lattice
    assert "Edge Count: 395" in log and "Max Fitness: 756" in log and "Starting Fitness: 0.52248675" in log
    final = float(re.findall(r"\nFitness: ([0-9.]+)", log)[-1])
    assert 0.9 < final <= 1.0                                 # the evolved genome aggregates (0.987 +- 0.004)
    # the standalone run batches its trials: 3 sizes x 5 seeds = 15 trials, one launch per snapshot stage (initial
    # configuration, after step n, after step n^2, final) instead of one per trial and stage
    r = subprocess.run([exe, "-b", "agg", "-e", "gm", "-k(6,4)", "-k(10,7)", "-k(13,9)", "-s", "1", "-s", "2", "-s", "3", "-s", "4", "-s", "62813",
                        "--gran", "10", "--path", str(gfile), "--out", str(out_dir), "--ga-seed", "9", "-v"],
                       capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    m = re.search(r'Please check: "(Agg_GM_3_sizes_5_trials_gran_10_\d+)" file', r.stdout)
    log = open(os.path.join(str(out_dir), m.group(1) + ".log")).read()
    assert int(re.search(r"Kernel launches for 15 trials: (\d+)", log).group(1)) <= 4
    assert len(re.findall(r"SOPS grid\n", log)) == 15 * 4 and log.count("Trial Elapsed Time:") == 15
    grids27 = re.findall(r"SOPS grid\n((?: [0-9 ]+\n){27})(?! )", log)
    last_trial_first = np.asarray([[int(v) for v in row.split()] for row in grids27[-4].strip("\n").split("\n")], dtype=np.uint8)
    assert np.array_equal(last_trial_first, np.asarray(fixtures["agg_13_9_snapshots"][0], dtype=np.uint8))   # (13,9) / 62813, step 0
    # a short GA: 3 generations, population 6, two sizes x two seeds
    r = subprocess.run([exe, "-b", "sep", "-e", "ga", "-k", "(6,3)", "-k", "(7,4)", "-s", "1", "-s", "2", "--gran", "10", "-p", "6", "-g", "3",
                        "--out", str(out_dir), "--ga-seed", "4"], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    m = re.search(r'Please check: "(Sep_GA_2_sizes_2_trials_gran_10_\d+)" file', r.stdout)
    log = open(os.path.join(str(out_dir), m.group(1) + ".log")).read()
    assert log.count("Starting Gen:") == 3 and "Starting Separation GA Experiment..." in log
    avgs = [float(v) for v in re.findall(r"Avg. Fitness -> ([0-9.]+)", log)]
    assert all(0.2 < a < 1.0 for a in avgs)
    # invalid size (BASELINE's "(3,6)": more particles than cells) is an error, not a hang
    r = subprocess.run([exe, "-b", "agg", "-e", "gm", "-k(3,6)", "-s", "1", "--gran", "10", "--path", str(gfile), "--out", str(out_dir)],
                       capture_output=True, text=True, cwd=str(tmp_path), timeout=60)
    assert r.returncode == 1 and "invalid size" in r.stderr
