"""CPU: the parts of bench.py and integration/ that run without a GPU keep their contracts."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU leg: oracle port on all host threads) prints ONE JSON line with
    the keys the driver reads."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--cpu-rows", "300000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"].startswith("rows/sec groupby-sum")
    assert j["unit"] == "rows/s" and j["higher_is_better"] is True and j["value"] > 0
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
    assert j["e2e"]["value"] == j["value"] and j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    assert j["steps"] == 1 and "workload" in j["config"]


def test_hook_script_anchors_match_the_reference():
    """integration/apply_hook.py inserts next to short anchor strings: each must occur exactly once in the
    reference revision the goldens were generated from (skipped where /root/reference is absent), and
    the script must refuse to touch /root/reference itself."""
    sys.path.insert(0, os.path.join(ROOT, "integration"))
    try:
        import apply_hook as ah
    finally:
        sys.path.pop(0)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "integration", "apply_hook.py"), "/root/reference"],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "refusing" in (r.stdout + r.stderr)
    ref = "/root/reference"
    if not os.path.exists(os.path.join(ref, "src", "core", "sort.cc")):
        pytest.skip("no reference tree here")
    sort_cc = open(os.path.join(ref, "src", "core", "sort.cc")).read()
    red_cc = open(os.path.join(ref, "src", "core", "expr", "fexpr_reduce_unary.cc")).read()
    ext_py = open(os.path.join(ref, "ci", "ext.py")).read()
    for text, anchors in ((sort_cc, (ah.INCLUDE_ANCHOR, ah.OPTION_ANCHOR, ah.REGISTER_ANCHOR, ah.HOOK_ANCHOR)),
                          (red_cc, (ah.RED_INCLUDE_ANCHOR, ah.RED_HELPER_ANCHOR, ah.RED_LOOP_ANCHOR)),
                          (ext_py, (ah.EXT_ANCHOR,))):
        for a in anchors:
            assert text.count(a) == 1, a
