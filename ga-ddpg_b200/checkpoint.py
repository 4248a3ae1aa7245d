"""Reference-format checkpoints for the fused agents.

The reference saves three files per agent (core/agent.py:282-352):
    {dir}/{name}_actor_{env}_{surfix}       {"net": policy.state_dict(), "opt": Adam.state_dict(), "sch": MultiStepLR.state_dict()}
    {dir}/{name}_critic_{env}_{surfix}      same for the critic
    {dir}/{name}_state_feat_{env}_{surfix}  {"net", "opt", "encoder_opt", "sch", "encoder_sch", "val_encoder_opt",
                                             "val_encoder_sch", "step"}
and reads them back in ``load_model`` (agent.py:354-431).  Here Adam runs fused on flat arenas
``[params | grads | m | v]`` (nets.Arena), so the torch ``Adam.state_dict()`` layout is produced from / scattered
into the arena's moment ranges.  Everything in this file is host-side bookkeeping on tensors it is handed (it never
launches a kernel), which is why it can be unit-tested against a real ``torch.optim.Adam`` without a GPU.

torch layout (torch/optim/optimizer.py ``state_dict``): ``{"state": {i: {"step", "exp_avg", "exp_avg_sq"}},
"param_groups": [{..hyper-parameters.., "params": [0..n-1]}]}`` with ``i`` the position of the parameter in the
order it was handed to the optimiser (``module.parameters()``).  Parameters that never received a gradient have no
``state`` entry (``log_std_linear`` of the policy, ``extra_pred`` when the aux loss is off).
"""
import os

import torch


def paths(output_dir, name, env_name, surfix):
    """File names of agent.py:297-308 / 361-372."""
    mk = lambda part: "{}/{}_{}_{}_{}".format(output_dir, name, part, env_name, surfix)  # noqa: E731
    return {"actor": mk("actor"), "critic": mk("critic"), "goal_feat": mk("goal_feat"), "state_feat": mk("state_feat")}


def _param_groups(params, lr, eps, weight_decay, initial_lr=None):
    """Let torch itself write the hyper-parameter block so that its key set matches the installed version."""
    shadow = [torch.nn.Parameter(torch.empty(0)) for _ in params]
    opt = torch.optim.Adam(shadow, lr=float(lr), eps=eps, weight_decay=weight_decay)
    groups = opt.state_dict()["param_groups"]
    groups[0]["initial_lr"] = float(lr if initial_lr is None else initial_lr)
    return groups


def adam_state_dict(params, moments, step, lr, eps, weight_decay, initial_lr=None):
    """``params``: parameters in optimiser order; ``moments(i, p)`` -> (exp_avg, exp_avg_sq) views or None when the
    fused Adam never touches that parameter; ``step``: number of Adam steps taken so far (0 => empty state, like
    a freshly built torch optimiser)."""
    state = {}
    if step > 0:
        for i, p in enumerate(params):
            mv = moments(i, p)
            if mv is None:
                continue
            m, v = mv
            state[i] = {"step": torch.tensor(float(step)), "exp_avg": m.detach().clone().view(p.shape),
                        "exp_avg_sq": v.detach().clone().view(p.shape)}
    return {"state": state, "param_groups": _param_groups(params, lr, eps, weight_decay, initial_lr)}


def load_adam_state_dict(sd, params, moments):
    """Scatter a torch ``Adam.state_dict()`` into the arena moment views.  Returns (step, lr).  Parameters without
    a state entry get zero moments (what torch would lazily create).  Accepts the int ``step`` of old torch
    versions and the tensor ``step`` of current ones.  Shape mismatches raise (the caller mirrors the reference's
    try/except around the feature-extractor optimisers, agent.py:410-419)."""
    groups = sd["param_groups"]
    order = [i for g in groups for i in g["params"]]
    if len(order) != len(params):
        raise ValueError("optimizer state has %d parameters, the network has %d" % (len(order), len(params)))
    step = 0
    for pos, p in enumerate(params):
        mv = moments(pos, p)
        st = sd["state"].get(order[pos])
        if st is None:
            if mv is not None:
                mv[0].zero_(), mv[1].zero_()
            continue
        if tuple(st["exp_avg"].shape) != tuple(p.shape):
            raise ValueError("exp_avg shape %s does not match parameter shape %s" % (tuple(st["exp_avg"].shape), tuple(p.shape)))
        if mv is None:
            continue  # state for a parameter the fused step never updates: nothing to keep
        mv[0].view(p.shape).copy_(st["exp_avg"])
        mv[1].view(p.shape).copy_(st["exp_avg_sq"])
        step = max(step, int(float(st["step"])))
    return step, float(groups[0]["lr"])


def arena_moments(arena, ranges=None):
    """moments() callback for parameters whose ``.data`` are views into ``arena.p`` (nets.Arena).  ``ranges``:
    optional list of (offset, count) float ranges the fused Adam steps; parameters outside them return None."""
    base, esz = arena.p.data_ptr(), arena.p.element_size()

    def moments(i, p):
        off = (p.data_ptr() - base) // esz
        n = p.numel()
        if off < 0 or off + n > arena.n:
            raise ValueError("parameter %d is not a view into this arena" % i)
        if ranges is not None and not any(o <= off and off + n <= o + c for o, c in ranges):
            return None
        return arena.m[off: off + n], arena.v[off: off + n]

    return moments


def save(obj, path):
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    torch.save(obj, path)


def load(path, map_location="cpu", trusted=None):
    """Read one checkpoint file.  The reference calls plain ``torch.load`` (agent.py:374), i.e. full pickle; here the
    default is torch's restricted unpickler (tensors, containers, numbers) plus ``collections.Counter`` — the only extra
    type a ``MultiStepLR.state_dict()`` holds — which reads every file the reference or this package writes.  Files that
    need arbitrary pickle (e.g. written by very old torch versions) load only with ``trusted=True`` or
    ``GADDPG_TRUSTED_CHECKPOINTS=1``: unpickling executes code, so only do that for files you produced yourself."""
    import collections
    import pickle

    if trusted is None:
        trusted = os.environ.get("GADDPG_TRUSTED_CHECKPOINTS", "0") == "1"
    if trusted:
        return torch.load(path, map_location=map_location, weights_only=False)
    try:
        with torch.serialization.safe_globals([collections.Counter]):
            return torch.load(path, map_location=map_location, weights_only=True)
    except pickle.UnpicklingError as e:
        raise RuntimeError("%s needs full pickle to load (%s); pass trusted=True / set GADDPG_TRUSTED_CHECKPOINTS=1 if you "
                           "trust its origin" % (path, str(e).splitlines()[0])) from e


def migrate_model(in_model, out_model, surfix="latest", grasp_model=None, env_name="PandaYCBEnv"):
    """/root/reference/core/utils.py:319-334: copy a (BC or DDPG) checkpoint's files to DDPG names in another directory
    — how a behaviour-cloning run is turned into the initialisation of a DDPG run.  Same file set and the same fallback
    (``BC_*`` if present, else ``DDPG_*``, sticky once switched, like the reference's loop); returns the copies made."""
    import shutil

    in_policy, out_policy, done = "BC", "DDPG", []
    os.makedirs(out_model, exist_ok=True)
    for part in ("actor", "state_feat", "goal_feat", "critic"):
        name = "{}_{}_{}".format(part, env_name, surfix)
        if not os.path.exists("{}/{}_{}".format(in_model, in_policy, name)):
            in_policy = "DDPG"
        src, dst = "{}/{}_{}".format(in_model, in_policy, name), "{}/{}_{}".format(out_model, out_policy, name)
        if os.path.exists(src):
            if os.path.abspath(src) != os.path.abspath(dst):
                shutil.copyfile(src, dst)
            done.append((src, dst))
    return done
