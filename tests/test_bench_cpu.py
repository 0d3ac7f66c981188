"""Host-side pieces of bench.py that need no GPU: the roofline denominator and the JSON contract keys."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _with_peaks(tmp_path, monkeypatch, content):
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    if content is not None:
        (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps(content))
    return bench.measured_peak()


def test_fallback_peak_when_file_absent(tmp_path, monkeypatch):
    v, src = _with_peaks(tmp_path, monkeypatch, None)
    assert v == bench.FALLBACK_HBM_GBS and src.startswith("fallback")


def test_measured_peak_reads_hbm_gbs(tmp_path, monkeypatch):
    v, src = _with_peaks(tmp_path, monkeypatch, {"hbm_gbs": 6551.0, "bf16_tflops": 1686.7})
    assert v == 6551.0 and src.startswith("measured")


def test_measured_peak_falls_back_on_a_file_without_the_key(tmp_path, monkeypatch):
    assert _with_peaks(tmp_path, monkeypatch, {"bf16_tflops": 1800})[0] == bench.FALLBACK_HBM_GBS
    (tmp_path / "MEASURED_PEAKS.json").write_text("not json")
    assert bench.measured_peak()[0] == bench.FALLBACK_HBM_GBS


def test_committed_bench_line_has_the_contract_keys():
    with open(os.path.join(ROOT, "profiles", "r01_final_bench.json")) as f:
        d = json.loads(f.read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"])
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"])
