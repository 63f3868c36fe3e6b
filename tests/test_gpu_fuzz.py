"""Randomised GPU-vs-oracle sweeps (tests/fuzz_parity.py, tests/fuzz_synth.py) at sizes that finish in seconds: random mechanism / law /
batch / steps / kernel variant / record / resumed-call combinations of the rollout path, and random (n, nu, m_lambda, N, batch) Riccati
recursions across the kernel-selection boundaries.  Tolerance 1e-9 (north_star); failure flags, break indices and iteration counts exact.
The 200- and 150-case runs of round 2 are on file in profiles/fuzz_r2.jsonl and profiles/fuzz_synth_r2.jsonl."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [21, 22])
def test_rollout_random_combinations(cc, orc, seed):
    import fuzz_parity
    lines = []
    s = fuzz_parity.run(24, seed, emit=lines.append)
    assert s["errors"] == 0 and s["cases_above_1e9"] == 0, [l for l in lines if '"error": null' not in l or "false" in l][:4]
    assert s["worst_rel_err"] < 1e-9


@pytest.mark.parametrize("seed", [31, 32])
def test_riccati_random_shapes(cc, orc, seed):
    import fuzz_synth
    lines = []
    s = fuzz_synth.run(40, seed, emit=lines.append)
    assert s["cases_failing"] == 0, [l for l in lines if "false" in l or '"error": null' not in l][:4]
    assert s["worst_gain_err"] < 1e-9
