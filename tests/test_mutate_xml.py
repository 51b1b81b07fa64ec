"""hsr_env_b200.mjcf.mutate_tree == the reference's mutate_xml (/root/reference/hsr/util.py:87-182): same surviving joints
and actuators, same block bodies, same --set-xml edits, for the README model (C1), +1 block (C1b), the full arm (C3),
four blocks (C5), a --set-xml run and the cupboard scene.

The committed fixture tests/golden/mutate.json holds sha256 digests of the trees the REFERENCE'S OWN function produced
(tests/golden/make_mutate_golden.py runs it under a stubbed gym); where /root/reference is present (this container) the
reference is also run live and the full canonical trees are compared.  The CLI line of the README then goes through
add_env_args / add_wrapper_args / hierarchical_parse_args / env_wrapper exactly as /root/reference/hsr/control.py:78-86
does."""
import hashlib
import json
import sys
import xml.etree.ElementTree as ET
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN

sys.path.insert(0, str(GOLDEN))
import make_mutate_golden as ref_side  # noqa: E402  (case table, canonical form, the stubbed import of the reference)

from hsr_env_b200 import mjcf  # noqa: E402

ASSETS = mjcf.default_assets_root()
FIXTURE = json.loads((GOLDEN / "mutate.json").read_text())


def our_mutation(case):
    """The same mutation with hsr_env_b200.mjcf: main file + its includes, each mutated as the reference does (util.py:171-175)."""
    xml, dofs, n_blocks, points, changes = ref_side.CASES[case]
    main, included = mjcf.load_trees(Path(ASSETS, xml))
    setters = [mjcf.XMLSetter(p, v) for p, v in changes]
    block_pos = [points[i % len(points)] for i in range(n_blocks)]
    out = {}
    # the reference mutates the included files first, then the main file, drawing goal_space.sample() for every tree that
    # has a non-empty worldbody; only the main file has one
    for k, (name, tree) in enumerate(included.items()):
        mjcf.mutate_tree(tree, dofs, n_blocks, [np.asarray(p, float) for p in block_pos], setters)
        out[f"include{k}"] = ref_side.canonical(tree.getroot(), ref_side.DROP)
    mjcf.mutate_tree(main, dofs, n_blocks, [np.asarray(p, float) for p in block_pos], setters)
    out = {"main": ref_side.canonical(main.getroot(), ref_side.DROP), **out}
    return out


needs_assets = pytest.mark.skipif(ASSETS is None, reason="HSR assets (hsr/models) not found: set HSR_ASSETS")


@needs_assets
@pytest.mark.parametrize("case", list(ref_side.CASES))
def test_mutate_tree_matches_reference_digest(case):
    ours = our_mutation(case)
    want = FIXTURE[case]
    got = ref_side.digest(ours)
    assert got["joints"] == want["joints"]
    assert got["actuators"] == want["actuators"]
    assert [list(b) for b in got["blocks"]] == [list(b) for b in want["blocks"]]
    for k in ours:
        assert got[k] == want[k], f"{case}/{k}: mutated tree differs from the reference's"


@pytest.mark.skipif(not ref_side.REF.exists(), reason="the reference tree is only present in the build container")
@pytest.mark.parametrize("case", list(ref_side.CASES))
def test_mutate_tree_matches_reference_live(case):
    mod = ref_side.load_reference_util()
    want = ref_side.reference_mutation(mod, case)
    ours = our_mutation(case)
    assert set(ours) == set(want)
    for k in want:
        assert ours[k] == want[k]
    for k in want:
        assert hashlib.sha256(want[k].encode()).hexdigest() == FIXTURE[case][k], "tests/golden/mutate.json is stale"


@needs_assets
def test_readme_command_line_through_env_wrapper():
    """README.md:5 of the reference, parsed and wrapped as hsr/control.py:78-86 does; the wrapped main receives env_args
    with the compiled model, goals=[GoalSpec(block_space, goal_space, geofence)] and starts={} (hsr/util.py:69-74)."""
    import argparse

    from hsr_env_b200 import util
    from hsr_env_b200.model import Model

    argv = ("--block-space (0,0)(0,0)(0,0)(0,0) --goal-space (0,0)(0,0)(0,0) --use-dof slide_x --use-dof slide_y "
            "--steps-per-action=300 --geofence=.5 --n-blocks 1").split()
    parser = argparse.ArgumentParser()
    wrapper_parser = parser.add_argument_group("wrapper_args")
    env_parser = parser.add_argument_group("env_args")
    util.add_env_args(env_parser)
    util.add_wrapper_args(wrapper_parser)
    args = util.hierarchical_parse_args(parser, argv=argv)
    assert set(args) == {"wrapper_args", "env_args"}
    assert args["wrapper_args"]["use_dof"] == ["slide_x", "slide_y"] and args["wrapper_args"]["geofence"] == .5
    assert args["env_args"]["steps_per_action"] == 300
    seen = {}

    def main(env_args):
        seen.update(env_args)
        return Model.load(env_args["xml_file"]) if str(env_args["xml_file"]).endswith(".hsrb") else env_args["xml_file"]

    model = util.env_wrapper(main)(**args)
    assert seen["starts"] == {} and len(seen["goals"]) == 1
    g = seen["goals"][0]
    assert g.distance == .5 and g.a.shape == (4,) and g.b.shape == (3,)
    assert isinstance(model, Model) and (model.nq, model.nv, model.nu) == (9, 8, 2)      # C1b
    # ... and it is the model the committed blob holds: the GPU test of the single-environment facade
    # (tests/test_gpu_parity.py::test_control_loop_through_the_single_env_facade) starts from that blob, because the
    # reference's model files do not travel to the GPU box
    from conftest import BLOBS
    assert model.to_blob() == Model.load(BLOBS / "c1b_readme_block.hsrb").to_blob()
