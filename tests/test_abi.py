"""CPU-only: the C-ABI library builds, loads and exports every symbol include/ccx.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "ccx.h")).read()
    return sorted(set(re.findall(r"CCX_API\s+[\w\s\*]+?\b(ccx_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from imagecaptioningconvnext_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libccx.so missing: run `make`"
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ccx.h but not exported"


def test_python_binding_covers_header():
    from imagecaptioningconvnext_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_version_and_status_strings_without_gpu():
    from imagecaptioningconvnext_b200 import _lib
    L = _lib.lib()
    assert L.ccx_version() >= 100
    assert L.ccx_status_string(0) == b"ok"
    assert b"shape" in L.ccx_status_string(-1)


def test_struct_layouts_match_header_sizes():
    # sizes the C compiler sees (compiled on the fly with gcc) must equal the ctypes mirrors
    import subprocess
    import tempfile
    from imagecaptioningconvnext_b200 import _lib
    code = ('#include "ccx.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(ccx_linear_desc),'
            'sizeof(ccx_cnblock_weights), sizeof(ccx_downsample_weights), sizeof(ccx_encoder_weights));return 0;}\n')
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(code)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(_lib.LinearDesc), ctypes.sizeof(_lib.CNBlockWeights),
                     ctypes.sizeof(_lib.DownsampleWeights), ctypes.sizeof(_lib.EncoderWeights)]


def test_encoder_state_dict_keys_match_torchvision():
    import torchvision
    from imagecaptioningconvnext_b200 import Encoder
    ref = torchvision.models.convnext_base(weights=None).features.state_dict()
    sd = Encoder().state_dict()
    assert list(sd) == ["convnext." + k for k in ref]
    assert all(sd["convnext." + k].shape == v.shape for k, v in ref.items())


def test_encoder_refuses_cpu_tensors():
    import pytest
    import torch
    from imagecaptioningconvnext_b200 import Encoder
    with pytest.raises(ValueError):
        Encoder()(torch.zeros(1, 3, 64, 64))


def test_integration_md_binding_example_matches_the_abi():
    """The ctypes struct shown to maintainers in INTEGRATION.md is the one the library expects."""
    import ctypes
    import os
    import re
    from imagecaptioningconvnext_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "INTEGRATION.md")).read()
    m = re.search(r"class LinearDesc\(ctypes.Structure\):.*?\n\n", src, re.S)
    assert m, "INTEGRATION.md lost its LinearDesc example"
    ns = {"ctypes": ctypes}
    exec(m.group(0), ns)
    doc = ns["LinearDesc"]
    assert [f[0] for f in doc._fields_] == [f[0] for f in _lib.LinearDesc._fields_]
    assert ctypes.sizeof(doc) == ctypes.sizeof(_lib.LinearDesc)
