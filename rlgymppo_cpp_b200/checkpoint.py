"""Checkpoints in the reference's on-disk layout, so runs move between RLGymPPO_CPP and this engine.

Reference: /root/reference/RLGymPPO_CPP/src/public/RLGymPPO_CPP/Learner.cpp:171-309 (SaveStats / LoadStats / Save / Load),
src/private/RLGymPPO_CPP/PPO/PPOLearner.cpp:362-502 (TorchLoadSaveSeq / TorchLoadSaveAll), tools/checkpoint_converter.py.

  <checkpointSaveFolder>/<cumulative timesteps>/
      PPO_POLICY.lt, PPO_CRITIC.lt      torch::save(nn::Sequential) == a TorchScript archive of Linear/ReLU modules whose
                                        parameters are named "0.weight", "0.bias", "2.weight", ...  Written here with
                                        torch.jit.script(seq).save (what the reference's own converter feeds to torch::load),
                                        read with torch.jit.load.
      PPO_POLICY_OPTIM.lt, PPO_CRITIC_OPTIM.lt
                                        torch::save(optim::Adam) is a libtorch-only archive: the reference's converter cannot
                                        translate it either ("optimizers will be reset").  This engine stores its Adam state
                                        next to them as *_OPTIM.pt (torch.save) and resets the optimiser when only the .lt
                                        file exists; the reference ignores the .pt files and resets too
                                        (TorchLoadSaveAll: missing optimizer files are not fatal, PPOLearner.cpp:430-470).
      RUNNING_STATS.json                cumulative_timesteps, cumulative_model_updates, epoch, skill_rating,
                                        reward_running_stats {mean, var, shape, count}, run_id (Learner.cpp:171-242)
"""
from __future__ import annotations

import json
import os
import shutil
from typing import Optional

import torch

STATS_FILE_NAME = "RUNNING_STATS.json"
MODEL_FILE_NAMES = ("PPO_POLICY.lt", "PPO_CRITIC.lt")          # PPOLearner.cpp:362-365
OPTIM_FILE_NAMES = ("PPO_POLICY_OPTIM.lt", "PPO_CRITIC_OPTIM.lt")
OPTIM_PT_NAMES = ("PPO_POLICY_OPTIM.pt", "PPO_CRITIC_OPTIM.pt")


def seq_sizes(seq) -> list:
    """GetSeqSizes: element counts of every parameter, in order (PPOLearner.cpp:352-360)."""
    return [int(p.numel()) for p in seq.parameters()]


def save_seq(seq: torch.nn.Sequential, path: str):
    cpu = torch.nn.Sequential(*[type(m)(m.in_features, m.out_features) if isinstance(m, torch.nn.Linear) else type(m)() for m in seq])
    cpu.load_state_dict({k: v.detach().cpu() for k, v in seq.state_dict().items()})
    torch.jit.script(cpu).save(path)


def load_seq(seq: torch.nn.Sequential, path: str):
    if not os.path.isfile(path):
        raise RuntimeError(f"Failed to load from {path}, file does not exist or can't be accessed")
    try:
        loaded = torch.jit.load(path, map_location="cpu").state_dict()
    except Exception as ex:  # noqa: BLE001
        raise RuntimeError(f"Failed to load model, checkpoint may be corrupt or of different model arch.\nException: {ex}")
    before = seq_sizes(seq)
    after = [int(v.numel()) for k, v in loaded.items()]
    if before != after:  # "Torch will happily load in a model of a totally different size" (PPOLearner.cpp:390-410)
        raise RuntimeError(f"Saved model has different size than current model, cannot load model from {path}:\n"
                           f" > Current model: [ {' '.join(map(str, before))}  ],\n > Saved model:   [ {' '.join(map(str, after))}  ]")
    own = seq.state_dict()
    if set(own.keys()) != set(loaded.keys()):
        raise RuntimeError(f"Saved model has different parameter names than current model, cannot load model from {path}")
    with torch.no_grad():
        for k, v in loaded.items():
            own[k].copy_(v.reshape(own[k].shape))


def save_stats(path: str, total_timesteps: int, cumulative_model_updates: int, total_epochs: int, return_stats, skill_rating=None, run_id=None):
    j = {"cumulative_timesteps": int(total_timesteps), "cumulative_model_updates": int(cumulative_model_updates), "epoch": int(total_epochs)}
    if skill_rating is not None:
        j["skill_rating"] = skill_rating
    # WelfordRunningStat of shape 1: running mean / variance vectors (WelfordRunningStat.h:36-83)
    j["reward_running_stats"] = {"mean": [float(return_stats.mean)], "var": [float(return_stats.var)], "shape": 1, "count": int(return_stats.count)}
    if run_id is not None:
        j["run_id"] = run_id
    with open(path, "w") as f:
        f.write(json.dumps(j, indent=4))


def load_stats(path: str) -> dict:
    if not os.path.isfile(path):
        raise RuntimeError(f"Learner::LoadStats(): Can't open file at {path}")
    with open(path) as f:
        return json.load(f)


def numbered_folders(folder: str):
    out = []
    if os.path.isdir(folder):
        for name in os.listdir(folder):
            if os.path.isdir(os.path.join(folder, name)):
                try:
                    out.append(int(name))
                except ValueError:
                    pass
    return sorted(out)


def save_learner(learner, folder: Optional[str] = None) -> str:
    """Learner::Save (Learner.cpp:244-281)."""
    cfg = learner.cfg
    folder = folder if folder is not None else cfg.checkpointSaveFolder
    if not folder:
        raise RuntimeError("Learner::Save(): Cannot save because config.checkpointSaveFolder is not set")
    dst = os.path.join(folder, str(int(learner.total_timesteps)))
    os.makedirs(dst, exist_ok=True)
    st = learner.skill_tracker
    rating = None
    if st is not None:
        rating = dict(st.cur_rating) if st.config.perModeRatings else st.cur_rating[""]
    save_stats(os.path.join(dst, STATS_FILE_NAME), learner.total_timesteps, learner.ppo.cumulative_model_updates, learner.total_epochs,
               learner.return_stats, rating)
    for seq, name in zip((learner.ppo.policy, learner.ppo.value_net), MODEL_FILE_NAMES):
        save_seq(seq, os.path.join(dst, name))
    for opt, name in zip((learner.ppo.policy_opt, learner.ppo.value_opt), OPTIM_PT_NAMES):
        torch.save(opt.state_dict(), os.path.join(dst, name))
    if cfg.checkpointsToKeep != -1:  # remove the oldest (counted in the LOAD folder, like the reference: Learner.cpp:256-278)
        nums = numbered_folders(cfg.checkpointLoadFolder)
        if len(nums) > cfg.checkpointsToKeep:
            shutil.rmtree(os.path.join(cfg.checkpointLoadFolder, str(nums[0])), ignore_errors=False)
    return dst


def load_learner(learner, folder: Optional[str] = None) -> Optional[str]:
    """Learner::Load (Learner.cpp:283-365): the highest-numbered sub-folder; None when there is no checkpoint."""
    cfg = learner.cfg
    folder = folder if folder is not None else cfg.checkpointLoadFolder
    if not folder:
        raise RuntimeError("Learner::Load(): Cannot load because config.checkpointLoadFolder is not set")
    nums = numbered_folders(folder)
    if not nums:
        return None
    src = os.path.join(folder, str(nums[-1]))
    j = load_stats(os.path.join(src, STATS_FILE_NAME))
    learner.total_timesteps = int(j["cumulative_timesteps"])
    learner.ppo.cumulative_model_updates = int(j["cumulative_model_updates"])
    learner.total_epochs = int(j["epoch"])
    rrs = j["reward_running_stats"]
    learner.return_stats.mean = float(rrs["mean"][0])
    learner.return_stats.var = float(rrs["var"][0])
    learner.return_stats.count = int(rrs["count"])
    st = learner.skill_tracker
    if st is not None and "skill_rating" in j:  # LoadRatingSet (SkillTracker.cpp:259-291)
        r = j["skill_rating"]
        if isinstance(r, dict):
            if st.config.perModeRatings:
                st.cur_rating = {st.mode: float(r.get(st.mode, st.config.initialRating))}
            else:
                st.cur_rating = {"": float(st.config.initialRating)}
        else:
            st.cur_rating = {st.mode: float(r)}
    for seq, name in zip((learner.ppo.policy, learner.ppo.value_net), MODEL_FILE_NAMES):
        load_seq(seq, os.path.join(src, name))
    for opt, name in zip((learner.ppo.policy_opt, learner.ppo.value_opt), OPTIM_PT_NAMES):
        p = os.path.join(src, name)
        if os.path.isfile(p):
            opt.load_state_dict(torch.load(p, map_location=learner.device))
    if hasattr(learner.ppo, "upload_modules"):  # host mirrors -> the device learner's parameters
        learner.ppo.upload_modules()
    learner._push_weights()
    if st is not None and st.config.loadOldVersionsFromCheckpoints:
        load_old_versions(learner, folder)
    return src


def load_rating_set(st, r) -> dict:
    """SkillTracker::LoadRatingSet (SkillTracker.cpp:259-291)."""
    if isinstance(r, dict):
        if st.config.perModeRatings:
            out = {st.mode: float(st.config.initialRating)}
            out.update({k: float(v) for k, v in r.items()})
            return out
        return {"": float(st.config.initialRating)}
    return {st.mode: float(r)}


def load_old_versions(learner, folder: str) -> int:
    """Learner::Load's old-version scan (Learner.cpp:311-371): for i = 1..maxVersions the checkpoint closest to
    totalTimesteps - i * timestepsPerVersion (below the previous target, at most one interval short) that carries a
    skill_rating becomes an old policy version with that rating."""
    st = learner.skill_tracker
    interval = int(st.config.timestepsPerVersion)
    target = int(learner.total_timesteps)
    found = 0
    for _ in range(st.config.maxVersions):
        target -= interval
        best, best_rating = -1, None
        for n in numbered_folders(folder):
            if n < target + interval and (best == -1 or abs(n - target) < abs(best - target)):
                sp = os.path.join(folder, str(n), STATS_FILE_NAME)
                if os.path.isfile(sp):
                    j = load_stats(sp)
                    if "skill_rating" in j:
                        best, best_rating = n, j["skill_rating"]
        if best != -1 and best >= target - interval:
            mp = os.path.join(folder, str(best), MODEL_FILE_NAMES[0])
            if os.path.isfile(mp):
                from .learner import make_mlp, mlp_layers_numpy

                seq = make_mlp(learner.engine.obs_size, list(learner.cfg.ppo.policyLayerSizes), learner.engine.num_actions)
                load_seq(seq, mp)
                st.append_old_policy(mlp_layers_numpy(seq), load_rating_set(st, best_rating))
                found += 1
    return found
