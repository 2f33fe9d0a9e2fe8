"""Host logic: the arch restatement (deepfake_detection_b200/arch.py) reproduces the reference's
state_dict / named_parameters names, shapes and order (fixture minted from the reference itself by
oracle/mint_goldens.py::mint_state_keys)."""
import json
import os

import pytest

from deepfake_detection_b200.arch import SUPPORTED_ARCHS, get_spec, param_entries, state_entries


@pytest.mark.parametrize("arch", SUPPORTED_ARCHS)
def test_state_and_param_entries_match_reference(arch, golden_dir):
    ref = json.load(open(os.path.join(golden_dir, "state_keys.json")))[arch]
    spec = get_spec(arch, num_classes=2, in_chans=12 if arch == "efficientnet_deepfake_v4" else 3)
    mine = [[n, list(s)] for n, s, _ in state_entries(spec)]
    assert mine == ref["state"]
    minep = [[n, list(s)] for n, s, _ in param_entries(spec)]
    assert minep == ref["params"]
    n = 0
    for _, s, _ in param_entries(spec):
        k = 1
        for d in s:
            k *= d
        n += k
    assert n == ref["n_params"]
