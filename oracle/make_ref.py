#!/usr/bin/env python
"""Recipe for oracle/_ref: the UNMODIFIED reference classes of the hot path, staged so that they travel to the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY (nothing under rl_x_b200/ may import this or oracle/_ref).

The reference (nico-bohlinger/RL-X) is pure Python, so "building" it is copying the handful of modules the PPO / SAC
PyTorch path consists of out of /root/reference, byte for byte, into oracle/_ref/rl_x/... (git-ignored: the sources never
enter this repository's history; NOT gpurun-ignored: the directory is shipped with the snapshot because /root/reference does
not exist on the GPU box).  `bench.py --impl reference` then imports `rl_x.algorithms.ppo.pytorch.ppo.PPO` from there and
times it on the host cores (oracle/ref_arm.py).  Run in the build container:

    python oracle/make_ref.py          # also called by __graft_entry__.build() when /root/reference is present

A manifest with the sha256 of every staged file is written next to them so that a stale or edited copy is detectable.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
DST = os.path.join(HERE, "_ref")

FILES = [
    "rl_x/__init__.py",
    "rl_x/algorithms/__init__.py",
    "rl_x/algorithms/algorithm.py",
    "rl_x/algorithms/algorithm_manager.py",
    "rl_x/algorithms/deep_learning_framework_type.py",
    "rl_x/algorithms/ppo/__init__.py",
    "rl_x/algorithms/ppo/pytorch/__init__.py",
    "rl_x/algorithms/ppo/pytorch/batch.py",
    "rl_x/algorithms/ppo/pytorch/critic.py",
    "rl_x/algorithms/ppo/pytorch/default_config.py",
    "rl_x/algorithms/ppo/pytorch/general_properties.py",
    "rl_x/algorithms/ppo/pytorch/policy.py",
    "rl_x/algorithms/ppo/pytorch/ppo.py",
    "rl_x/algorithms/sac/__init__.py",
    "rl_x/algorithms/sac/pytorch/__init__.py",
    "rl_x/algorithms/sac/pytorch/critic.py",
    "rl_x/algorithms/sac/pytorch/default_config.py",
    "rl_x/algorithms/sac/pytorch/entropy_coefficient.py",
    "rl_x/algorithms/sac/pytorch/general_properties.py",
    "rl_x/algorithms/sac/pytorch/policy.py",
    "rl_x/algorithms/sac/pytorch/q_network.py",
    "rl_x/algorithms/sac/pytorch/replay_buffer.py",
    "rl_x/algorithms/sac/pytorch/sac.py",
    # FastSAC: staged for the checkpoint interop test (tests/test_fastsac_emulation.py), not timed anywhere
    "rl_x/algorithms/fastsac/__init__.py",
    "rl_x/algorithms/fastsac/pytorch/__init__.py",
    "rl_x/algorithms/fastsac/pytorch/critic.py",
    "rl_x/algorithms/fastsac/pytorch/default_config.py",
    "rl_x/algorithms/fastsac/pytorch/entropy_coefficient.py",
    "rl_x/algorithms/fastsac/pytorch/fastsac.py",
    "rl_x/algorithms/fastsac/pytorch/general_properties.py",
    "rl_x/algorithms/fastsac/pytorch/observation_normalizer.py",
    "rl_x/algorithms/fastsac/pytorch/policy.py",
    "rl_x/algorithms/fastsac/pytorch/q_network.py",
    "rl_x/algorithms/fastsac/pytorch/replay_buffer.py",
    "rl_x/environments/__init__.py",
    "rl_x/environments/action_space_type.py",
    "rl_x/environments/data_interface_type.py",
    "rl_x/environments/observation_space_type.py",
    "rl_x/environments/simulation_type.py",
    "rl_x/environments/environment.py",
    "rl_x/environments/environment_manager.py",
]


def _sha(path):
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def available():
    """True when a staged copy is present (here after build(), on the GPU box via the snapshot)."""
    return os.path.exists(os.path.join(DST, "MANIFEST.json"))


def build_ref(ref_root=REF_ROOT):
    """Stage the reference modules. Returns the destination, or None when the reference tree is absent (GPU box)."""
    if not os.path.isdir(os.path.join(ref_root, "rl_x")):
        return DST if available() else None
    manifest = {}
    for rel in FILES:
        src = os.path.join(ref_root, rel)
        if not os.path.exists(src):
            if rel.endswith("__init__.py"):  # namespace-style package in the reference: an empty marker is enough
                os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
                open(os.path.join(DST, rel), "a").close()
                manifest[rel] = "(empty package marker, absent in the reference)"
                continue
            raise FileNotFoundError(src)
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    head = None
    try:
        with open(os.path.join(ref_root, ".git", "HEAD")) as fh:
            head = fh.read().strip()
    except OSError:
        pass
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": ref_root, "git_head": head, "files": manifest}, fh, indent=1, sort_keys=True)
    return DST


def verify():
    """sha256 of every staged file against the manifest (run on the GPU box before timing)."""
    with open(os.path.join(DST, "MANIFEST.json")) as fh:
        man = json.load(fh)
    bad = [rel for rel, h in man["files"].items() if len(h) == 64 and _sha(os.path.join(DST, rel)) != h]
    if bad:
        raise RuntimeError(f"oracle/_ref is not the staged reference any more: {bad}")
    return man


if __name__ == "__main__":
    out = build_ref(sys.argv[1] if len(sys.argv) > 1 else REF_ROOT)
    print(out or "reference tree not found; nothing staged")
