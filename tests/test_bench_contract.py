"""CPU-only: the bench.py contract on the arm that runs without a GPU.  `--impl reference` must print exactly ONE JSON
line on stdout (library banners go to stderr) carrying the keys the driver reads; the GPU arm refuses to run without a
device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--circuit", "hash"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "nova_fold_steps_per_sec" and d["unit"] == "steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_gpu_arm_refuses_without_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is visible")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
    assert p.stdout.strip() == ""


def test_bench_module_has_no_undefined_globals():
    """Every global name bench.py's functions load is defined in the module or is a builtin (a deleted helper would
    otherwise only surface on the multi-GPU leg, which runs nowhere but on the GPU box)."""
    import builtins
    import importlib.util
    import symtable
    path = os.path.join(ROOT, "bench.py")
    spec = importlib.util.spec_from_file_location("bench_under_test", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    missing = set()

    def walk(tab):
        for sym in tab.get_symbols():
            if sym.is_global() and sym.is_referenced() and not hasattr(mod, sym.get_name()) and not hasattr(builtins, sym.get_name()):
                missing.add((tab.get_name(), sym.get_name()))
        for child in tab.get_children():
            walk(child)

    walk(symtable.symtable(open(path).read(), path, "exec"))
    assert not missing, missing
