"""bench.py's reference arm runs on the host alone, so its JSON line -- the contract the driver parses -- is checked here:
one line on stdout, the required keys, a `config` that names the sample actually run, ranks other than 0 silent."""
import json
import os
import subprocess
import sys

from conftest import ROOT

REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
            "data", "config", "cpu_baseline", "e2e"}


def _run(env_extra=None, args=()):
    env = dict(os.environ)
    env.update(env_extra or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--alleles", "16", "--reads", "2000",
                        *args], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    return p.stdout.decode()


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = _run()
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert REQUIRED <= set(d), REQUIRED - set(d)
    assert d["impl"] == "reference" and d["unit"] == "records/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["e2e"] == {"value": d["value"], "unit": "records/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and "records" in cb["sample"]
    assert d["config"]["workload"].startswith("configs[1]") and "2000 x 150 bp" in d["config"]["workload"]  # the sample actually run
    assert "WHOLE sample" in cb["sample"] and set(d["seconds_by_phase"]) >= {"score", "select", "depthcap_pileup_consensus"}


def test_reference_arm_other_ranks_exit_silently():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}).strip() == ""
