"""The reference's own test files, run against this package (``oracle/ref_suite.py`` aliases ``pathpyG`` to
``pathpyg_b200``).  Without a GPU the tests that reach a kernel fail with the package's "needs a CUDA device" error --
those are tolerated here (they are covered by the ``-m gpu`` tier); any OTHER failure is a drop-in defect."""
import os
import re
import subprocess
import sys
import xml.etree.ElementTree as ET

import pytest

from oracle import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = os.path.join(ref_loader.REFERENCE_ROOT, "tests")
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="/root/reference not mounted")

# one pytest process per directory: the reference keeps a conftest.py with same-named fixtures in each of them
GROUPS = {
    "core": (["core"], 60),
    "io": (["io/test_pandas.py"], 52),
    "algorithms": (["algorithms/test_centrality.py", "algorithms/test_components.py", "algorithms/test_lift_order.py",
                    "algorithms/test_rolling_time_window.py", "algorithms/test_shortest_paths.py", "algorithms/test_temporal.py",
                    "algorithms/test_wl.py"], 10),
    "nn_utils": (["nn", "utils"], 1),
}


@pytest.mark.parametrize("group", sorted(GROUPS))
def test_reference_tests_pass_or_need_the_gpu(group, tmp_path):
    paths, min_passed = GROUPS[group]
    report = tmp_path / "report.xml"
    subprocess.run([sys.executable, "-m", "oracle.ref_suite", *[os.path.join(REF_TESTS, p) for p in paths], f"--junitxml={report}"],
                   cwd=ROOT, capture_output=True, text=True, timeout=600)
    cases = list(ET.parse(report).getroot().iter("testcase"))
    assert cases, "no reference test was collected"
    passed, needs_gpu, broken = 0, 0, []
    for case in cases:
        problems = [el for el in case if el.tag in ("failure", "error")]
        if not problems:
            passed += 1
        elif all(re.search("needs a CUDA device|CUDA", (el.get("message") or "") + (el.text or "")) for el in problems):
            needs_gpu += 1
        else:
            broken.append(f"{case.get('classname')}::{case.get('name')}: {(problems[0].get('message') or '')[:200]}")
    assert not broken, "reference tests failing for a reason other than the missing GPU:\n" + "\n".join(broken)
    assert passed >= min_passed, f"only {passed} reference tests passed ({needs_gpu} need the GPU)"
