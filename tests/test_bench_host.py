"""Host-side arithmetic of bench.py (no GPU, no model): the CPU arm's cost line and the per-config metric table."""
import numpy as np

import bench


def test_cost_line_recovers_base_and_marginal_cost():
    # cost(T) = base + T * marginal exactly (the reference schedule repeats both streams per tissue)
    pts = [(1, 5.0 + 3.0), (2, 5.0 + 6.0), (4, 5.0 + 12.0)]
    rate, base, marginal = bench.fit_rate(pts, stage1_s=0.5, T=63)
    assert abs(marginal - 3.0) < 1e-9 and abs(base - 5.5) < 1e-9
    assert abs(rate - 63 / (5.5 + 63 * 3.0)) < 1e-12


def test_cost_line_with_noise_is_least_squares_not_a_two_point_difference():
    rng = np.random.default_rng(0)
    pts = [(t, 4.0 + 2.5 * t + rng.normal(0, 0.05)) for t in (1, 2, 4, 1, 2, 4)]
    _, base, marginal = bench.fit_rate(pts, stage1_s=0.0, T=63)
    assert abs(marginal - 2.5) < 0.1 and abs(base - 4.0) < 0.3


def test_a_single_tissue_count_does_not_pretend_to_know_the_base_cost():
    # one point cannot separate base from marginal: the fallback attributes everything to the marginal cost (a LOWER CPU
    # rate) — which is why run_reference adds a warm-up sample at a second tissue count before it fits
    rate1, base, marginal = bench.fit_rate([(1, 8.0)], stage1_s=0.0, T=63)
    assert base == 0.0 and abs(marginal - 8.0) < 1e-12
    rate2, _, _ = bench.fit_rate([(1, 8.0), (2, 11.0)], stage1_s=0.0, T=63)
    assert rate2 > 2 * rate1


def test_every_config_has_a_metric_and_unit():
    for c in (1, 2, 3, 4, 5):
        metric, unit = bench.METRICS[c]
        assert metric and unit
    assert bench.METRICS[3][0] == bench.METRIC
