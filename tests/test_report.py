"""Host logic of the benchmark harness (gproshan_b200/report.py, §8 row f4): the reference's error measure, degree
histogram, topleset distribution and LaTeX row formats (src/test_geodesics_ptp.cpp). The GPU run of the harness is in
tests/test_gpu_variants.py::test_report_harness."""
import numpy as np

from gproshan_b200 import meshgen as mg, report


def test_compute_error_matches_reference_formula():
    exact = np.array([0.0, 1.0, 2.0, 4.0])
    dist = np.array([0.0, 1.1, 1.8, 4.0])
    # src/test_geodesics_ptp.cpp:361-370: 100 / (n - s) * sum_{exact > 0} |d - e| / e
    assert np.isclose(report.compute_error(dist, exact, 1), 100.0 / 3 * (0.1 / 1 + 0.2 / 2 + 0.0))


def test_degree_histogram_and_toplesets_distribution():
    g = mg.grid(5)            # 25 vertices: 9 interior (6 triangles), borders get +1 for the open fan
    h = report.degree_histogram(g)
    assert sum(h.values()) == 25 and h[6] == 9
    star = np.bincount(g.VT, minlength=25)
    border = np.array([g.EVT[v] != 0xFFFFFFFF and g.OT[g.EVT[v]] == 0xFFFFFFFF for v in range(25)])
    assert h == {int(d): int(c) for d, c in zip(*np.unique(star + border, return_counts=True))}
    sizes, ssorted = report.toplesets_distribution(np.array([0, 1, 7, 19, 25]))
    assert sizes.tolist() == [1, 6, 12, 6] and ssorted.tolist() == [1, 6, 6, 12]


def test_latex_rows():
    r = report.latex_row("bunny", 34834, 0.0123, 1.234, fm_s=0.5)
    assert r.startswith("        \\verb|bunny| &        34834 ") and r.endswith("\\\\\n")
    assert "&    \\bf  0.012s & \\bf (40.7x) &    \\bf   1.23\\% " in r   # "& %6s %6.3lfs " with "\\bf" (:36-40)
    d = report.latex_row("bunny", 34834, 0.0123, 1.234, double=True)
    assert d.count("\n") == 2 and "& Cuda " in d and d.endswith("\\\\\\hline\n") and "& OpenMP " in d


def test_analytic_exact():
    s = mg.icosphere(4)
    e = report.analytic_exact("sphere", s, 0)
    assert e[0] == 0 and np.isclose(e.max(), np.pi, atol=1e-6)
    g = mg.grid(3)
    assert np.isclose(report.analytic_exact("plane", g, 0)[8], np.sqrt(2.0))
