"""CLI flags + env wrapper of the reference, kept verbatim, feeding the batched backend.

NOTE: most of this file is the REFERENCE'S OWN CODE carried over as the interface schema (flag declarations, parser helper
bodies); it is not original work of this repository and sits off the hot path.

Mirrors /root/reference/hsr/util.py:16-81 (``add_env_args``, ``add_wrapper_args``, ``xml_setter``,
``env_wrapper``) and /root/reference/rl_utils/argparse.py:10-73 (``hierarchical_parse_args``, ``make_box``,
``parse_space``, ``parse_vector``).  Where the reference writes a mutated temp MJCF for MuJoCo to load
(util.py:161-179), ``env_wrapper`` here compiles the same mutation into a model blob (mjcf.compile_model) and
hands its path to the env as ``xml_file``.
"""
from __future__ import annotations

import argparse
import re
import tempfile
from collections import namedtuple
from functools import wraps
from pathlib import Path
from typing import List, Tuple

import numpy as np

from .spaces import Box

XMLSetter = namedtuple("XMLSetter", "path value")
GoalSpec = namedtuple("GoalSpec", "a b distance")  # /root/reference/hsr/env.py:20


# ----------------------------------------------------------------------------- rl_utils.argparse
def hierarchical_parse_args(parser: argparse.ArgumentParser, include_positional=False, argv=None):
    """{group title: {dest: value}, **ungrouped}.  Accepts both spellings of argparse's default group
    ('optional arguments' before Python 3.10, 'options' after; the reference pops only the former,
    rl_utils/argparse.py:43)."""
    args = parser.parse_args(argv)

    def key_value_pairs(group):
        for action in group._group_actions:
            if action.dest != "help":
                yield action.dest, getattr(args, action.dest, None)

    def get_positionals(groups):
        for group in groups:
            if group.title == "positional arguments":
                for _, v in key_value_pairs(group):
                    yield v

    def get_nonpositionals(groups):
        for group in groups:
            if group.title != "positional arguments":
                children = key_value_pairs(group)
                descendants = get_nonpositionals(group._action_groups)
                yield group.title, {**dict(children), **dict(descendants)}

    positional = list(get_positionals(parser._action_groups))
    nonpositional = dict(get_nonpositionals(parser._action_groups))
    optional = {}
    for title in ("optional arguments", "options"):
        optional.update(nonpositional.pop(title, {}))
    nonpositional = {**nonpositional, **optional}
    if include_positional:
        return positional, nonpositional
    return nonpositional


def make_box(*tuples: Tuple[float, float]):
    low, high = map(np.array, zip(*[tuple(map(float, m)) for m in tuples]))
    return Box(low=low, high=high, dtype=np.float32)


def parse_space(dim: int):
    def _parse_space(arg: str):
        pattern = r"\((-?[\.\d]+),(-?[\.\d]+)\)"
        regex = re.compile(pattern)
        matches = regex.findall(arg)
        if len(matches) != dim:
            raise argparse.ArgumentTypeError(
                f"Arg {arg} must have {dim} substrings matching pattern {regex.pattern}.")
        return make_box(*matches)

    return _parse_space


def parse_vector(length: int, delim: str):
    def _parse_vector(arg: str):
        vector = tuple(map(float, arg.split(delim)))
        if len(vector) != length:
            raise argparse.ArgumentTypeError(f'Arg {arg} must include {length} float values delimited by "{delim}".')
        return vector

    return _parse_vector


# ----------------------------------------------------------------------------- hsr.util
def add_env_args(parser):
    parser.add_argument("--obs-type", type=str, default=None)
    parser.add_argument("--render", action="store_true")
    parser.add_argument("--render-freq", type=int, default=None)
    parser.add_argument("--record", action="store_true")
    parser.add_argument("--record-freq", type=int, default=None)
    parser.add_argument("--record-path", type=Path, default=None)
    parser.add_argument("--steps-per-action", type=int, required=True)


def add_wrapper_args(parser):
    parser.add_argument("--block-space", type=parse_space(dim=4))
    parser.add_argument("--goal-space", type=parse_space(dim=3), required=True)
    parser.add_argument("--xml-file", type=Path, default="models/world.xml")
    parser.add_argument("--set-xml", type=xml_setter, action="append")
    parser.add_argument("--use-dof", type=str, action="append", default=[])
    parser.add_argument("--geofence", type=float, required=True)
    parser.add_argument("--n-blocks", type=int, default=0)


def add_batch_args(parser):
    """Flags the batched backend adds (none of them exist in the reference)."""
    parser.add_argument("--n-envs", type=int, default=1)
    parser.add_argument("--device", type=str, default="cuda:0")
    parser.add_argument("--seed", type=int, default=0)


def xml_setter(arg: str):
    return XMLSetter(*arg.split(","))


def env_wrapper(func):
    """Decorator with the reference's contract (util.py:53-81): consumes ``wrapper_args`` and calls
    ``func(env_args=...)`` with ``goals``, ``xml_file`` and ``starts`` filled in."""

    @wraps(func)
    def _wrapper(set_xml, use_dof, n_blocks, goal_space, xml_file, geofence, env_args: dict, block_space=None,
                 **kwargs):
        from . import mjcf
        from .env import get_xml_filepath

        xml_filepath = get_xml_filepath(xml_file)
        set_xml = list(set_xml or [])
        site_size = " ".join([str(geofence)] * 3)
        path = Path("worldbody", 'body[@name="goal"]', 'site[@name="goal"]', "size")
        set_xml += [XMLSetter(path=f"./{path}", value=site_size)]
        # blocks are spawned at goal_space.sample() at mutation time (util.py:107-108)
        block_pos = [goal_space.sample() for _ in range(n_blocks)]
        model = mjcf.compile_model(xml_filepath, use_dof, n_blocks=n_blocks, block_pos=block_pos,
                                   set_xml=[mjcf.XMLSetter(s.path, s.value) for s in set_xml], block_name="block")
        with tempfile.TemporaryDirectory() as tmp:
            temp_path = Path(tmp, xml_filepath.stem + ".hsrb")
            model.save_with_names(temp_path)
            env_args.update(
                goals=[GoalSpec(a=block_space, b=goal_space, distance=geofence)],
                xml_file=temp_path,
                starts={},
            )
            return func(env_args=env_args, **kwargs)

    def new_function(wrapper_args, **kwargs):
        return _wrapper(**wrapper_args, **kwargs)

    return new_function
