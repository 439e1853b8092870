"""Guard for the programmatic-dependent-launch (PDL) kernels: no read-only-path load above griddepcontrol.wait.

ptxas may move an ``ld.global.nc`` (SASS ``LDG.E.CONSTANT``: ``__ldg`` or a ``const __restrict__`` parameter) above the
wait (``ACQBULK``).  For a buffer the PDL primary writes that is a read of data that does not exist yet: it made the
whole-episode belief kernel read step-0 actions before the persistent rollout kernel had produced them.  Plain loads
above the wait are allowed -- they are the deliberate prologue reads of the static filter tables."""
import os
import re
import shutil
import subprocess

import pytest

from ia2c_b200 import _lib


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")
def test_no_readonly_path_load_above_the_pdl_wait():
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("library not built")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    offenders, kernels, name, seen_wait = [], 0, None, False
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name, seen_wait = m.group(1), False
            continue
        if "ACQBULK" in line:
            kernels += not seen_wait
            seen_wait = True
        elif name and not seen_wait and re.search(r"\bLDG\S*CONSTANT", line):
            offenders.append((name, line.strip()))
    # only kernels that HAVE a wait matter: drop the entries of kernels in which no ACQBULK follows
    with_wait = {n for n in re.findall(r"Function : (\S+)", sass)
                 if "ACQBULK" in sass.split("Function : " + n, 1)[1].split("Function : ", 1)[0]}
    offenders = [o for o in offenders if o[0] in with_wait]
    assert kernels > 10, "expected the PDL kernels to carry griddepcontrol.wait"
    assert not offenders, offenders[:5]
