"""Oracle tier T0: tests/t0_checklist.py run as a subprocess so that the real
``qdrant_client`` (if one is ever importable) is never shadowed by the test double on this process's sys.path.

  * real package present  -> every Appendix B item is checked against it: THIS is what pins the oracle;
  * absent (this image)   -> skipped, and the oracle stays "parity unpinned" (DESIGN.md §2);
  * the double            -> the checklist code itself is exercised (circular by construction; pins nothing)."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _run(*args):
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    return subprocess.run([sys.executable, os.path.join(HERE, "t0_checklist.py"), *args], capture_output=True,
                          text=True, env=env, timeout=600)


def test_checklist_runs_on_the_test_double():
    r = _run("--double")
    assert r.returncode == 0, r.stderr[-2000:]
    rep = json.loads(r.stdout.strip().splitlines()[-1])
    assert rep["pins_oracle"] is False and rep["client"] == "test double"
    assert len(rep["items"]) == 11
    assert rep["items"]["item6_root_filter_on_fusion"]["semantics"] == "legs"        # R4, as the engine does it
    assert rep["items"]["item9_duplicate_tie_order"]["order"] == [0, 2, 3]             # R5
    assert rep["items"]["random_differential"]["checked"] > 0


def test_oracle_against_real_qdrant_client():
    r = _run()
    if r.returncode == 77:
        pytest.skip("qdrant_client is not importable here: oracle stays PARITY UNPINNED (tier T0 not run)")
    assert r.returncode == 0, r.stderr[-4000:]
    rep = json.loads(r.stdout.strip().splitlines()[-1])
    assert rep["pins_oracle"] is True
    print("T0 report:", json.dumps(rep))


@pytest.mark.gpu
def test_t0_probe_on_gpu_pod():
    """run the T0 probe WHERE A WHEEL COULD EXIST (the GPU pod: site-packages, baseline/_ref,
    /opt/wheelhouse) and put the outcome on record.  Found -> the whole Appendix B checklist must pass against the real
    package (the oracle is then pinned).  Not found -> the exact lookup result is emitted as a warning (pytest prints
    warnings in its summary) and written to gpurun_out/t0_probe.json; the oracle stays PARITY UNPINNED."""
    import warnings
    r = _run("--probe")
    assert r.returncode == 0, r.stderr[-2000:]
    rep = json.loads(r.stdout.strip().splitlines()[-1])
    try:
        out_dir = os.path.join(os.path.dirname(HERE), "gpurun_out")
        os.makedirs(out_dir, exist_ok=True)
        json.dump(rep, open(os.path.join(out_dir, "t0_probe.json"), "w"), indent=1)
    except OSError:
        pass
    if not rep["found"]:
        warnings.warn(UserWarning("T0 probe on this box: qdrant_client NOT importable (" + rep["error"] + "); wheels found: "
                                  + json.dumps(rep["wheels_found"]) + "; python " + rep["python"]
                                  + "; oracle stays PARITY UNPINNED"))
        return
    full = _run()
    assert full.returncode == 0, full.stderr[-4000:]
    t0 = json.loads(full.stdout.strip().splitlines()[-1])
    assert t0["pins_oracle"] is True
    warnings.warn(UserWarning("T0 ran against qdrant-client " + rep["version"] + ": " + json.dumps(t0)))
