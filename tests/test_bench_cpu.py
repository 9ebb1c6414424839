"""bench.py's reference arm runs on the host cores only: its JSON line can be checked without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        *extra], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    return json.loads(lines[0])


def test_reference_arm_line_contract():
    d = _run()
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and base["metric"].startswith(d["metric"]) and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("cfg2") and d["config"]["mics_per_node"] == 4 and d["config"]["n_fft"] == 512
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    # the step time multiplies back to the frames the sample really processed
    frames = cb["value"] * d["ms_per_step"] / 1e3 * d["steps"]
    assert frames > 0 and abs(frames - round(frames / 626) * 626) < 1.0          # whole utterances of 626 frames


def test_reference_arm_other_workload():
    d = _run("--workload", "cfg3")
    assert d["config"]["workload"].startswith("cfg3") and d["config"]["nodes"] == 4


def test_numa_binding_helper_decodes_nvml_mask_and_survives_its_absence():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    before = os.sched_getaffinity(0)

    def broken(_):
        raise RuntimeError("no NVML")
    b._nvml_handle = broken
    assert b.bind_host_to_gpu(0) is None and os.sched_getaffinity(0) == before
    first = min(before)

    class FakeNvml:
        @staticmethod
        def nvmlDeviceGetCpuAffinity(handle, n_words):
            words = [0] * n_words
            words[first // 64] = 1 << (first % 64)
            return words
    b._nvml_handle = lambda i: (FakeNvml, None)
    try:
        if len(before) > 1:
            assert b.bind_host_to_gpu(0) == 1 and os.sched_getaffinity(0) == {first}
        else:
            assert b.bind_host_to_gpu(0) is None          # mask == allowed set: nothing to do
    finally:
        os.sched_setaffinity(0, before)
