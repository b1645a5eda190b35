"""The C++ host layer (include/rlgym_b200_shim.hpp) that mirrors the reference's plugin surface: the reference's example
app (examples/examplemain.cpp == T/examplemain.cpp's EnvCreateFunc + LearnerConfig) compiles against it with g++ and
links the C-ABI library.  Without a GPU it must fail LOUDLY (no CPU fallback); on the B200 it runs collection with the
built-in plugins and with a user-defined StateSetter (host path)."""
import json
import os
import subprocess
import sys

import pytest

from rlgymppo_cpp_b200 import build, meshes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def example_bin(tmp_path_factory):
    build.build()
    out = str(tmp_path_factory.mktemp("shim") / "examplemain")
    csrc = os.path.join(ROOT, "rlgymppo_cpp_b200", "csrc")
    env = dict(os.environ)
    env.pop("CXX", None)
    cmd = ["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-std=c++17", "-O2", "-Wall", "-Werror=return-type",
           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "examplemain.cpp"), "-o", out,
           "-L" + csrc, "-lrlgym_b200", "-Wl,-rpath," + csrc]
    subprocess.check_call(cmd, env=env)
    return out


@pytest.fixture(scope="module")
def mesh_dir(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("meshes"))
    meshes.write_placeholder_set(d)
    return d


def _has_gpu():
    import torch

    return torch.cuda.is_available()


def test_shim_compiles_links_and_fails_loudly_without_a_gpu(example_bin, mesh_dir):
    if _has_gpu():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([example_bin, mesh_dir, "1"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1
    assert "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr, r.stderr


def test_shim_rejects_missing_meshes(example_bin, tmp_path):
    r = subprocess.run([example_bin, str(tmp_path), "1"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "no collision meshes found" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("custom", [False, True, "mutators"])
def test_examplemain_collects_on_gpu(example_bin, mesh_dir, custom):
    args = [example_bin, mesh_dir, "3"] + (["--custom-setter"] if custom is True else ["--low-gravity"] if custom == "mutators" else [])
    r = subprocess.run(args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    last = json.loads(r.stdout.strip().splitlines()[-1])
    assert last["arenas"] == 16 * 24 and last["custom_setter"] == (custom is True)
    assert last["steps_per_second"] > 1000
    assert -5 < last["mean_step_reward"] < 5
    assert r.stdout.count("Timesteps Collected 100608") == 3  # ceil(100000 / 768) = 131 env-steps x 768 players
