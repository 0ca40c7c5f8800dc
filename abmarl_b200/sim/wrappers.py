"""Simulation wrappers -- mirror of abmarl/sim/wrappers/super_agent_wrapper.py over the batched engine.

`SuperAgentWrapper(sim, super_agent_mapping=...)` is the definition (same constructor, same assertions, same
`agents` dict: super agents + uncovered agents).  A manager built on it (abmarl_b200.managers) steps the inner sim's
learners exactly as before -- a super agent's action is the dict of its covered agents' actions, i.e. the same rows of
the action tensor -- and `SuperAgentView` re-reads the engine's per-learner outputs the way the wrapper does
(super_agent_wrapper.py:112-260):

    order   the wrapper unravels the super agents' actions first, in mapping order, then the uncovered agents'
            (:96-108): that is the order the inner sim processes them in, so the manager passes it to the engine
    obs     the covered agents' observations; a covered agent that was reported done in an EARLIER step shows its null
            observation (:153-160); mask[covered] = the covered agent is not done (:146-165)
    reward  the sum of the covered agents' rewards, a done agent's last reward counted once (:181-191)
    done    all covered agents are done (:211-216)
"""
import numpy as np
import torch

from abmarl_b200 import _capi as K
from abmarl_b200.spaces import Dict, Box
from abmarl_b200.sim import Agent


class SuperAgentWrapper:
    def __init__(self, sim, super_agent_mapping=None, **kwargs):
        self.sim = sim
        assert type(super_agent_mapping) is dict, "super agent mapping must be a dictionary."
        covered = set()
        for k, v in super_agent_mapping.items():
            assert type(k) is str, "The keys super agent mapping must be the super agent's id."
            assert k not in sim.agents, "A super agent cannot have the same id as an agent from the underlying sim."
            assert type(v) is list, "The values in super agent mapping must be lists of agent ids."
            for covered_agent in v:
                assert type(covered_agent) is str, "The covered agents list must be agent ids."
                assert covered_agent in sim.agents, "The covered agent must be an agent in the underlying sim."
                assert covered_agent not in covered, "The agent is already covered by another super agent."
                assert isinstance(sim.agents[covered_agent], Agent), "Covered agents must be learning Agents."
                covered.add(covered_agent)
        self.super_agent_mapping = super_agent_mapping
        self._covered_agents = covered
        self._uncovered_agents = [a for a in sim.agents if a not in covered]
        agents = {}
        for super_id, members in super_agent_mapping.items():          # :262-283
            obs = {'mask': Dict({m: Box(0, 1, (1,), int) for m in members})}
            obs.update({m: sim.agents[m].observation_space for m in members})
            agents[super_id] = Agent(id=super_id, observation_space=Dict(obs),
                                     action_space=Dict({m: sim.agents[m].action_space for m in members}))
        for agent_id in self._uncovered_agents:
            agents[agent_id] = sim.agents[agent_id]
        self.agents = agents

    @property
    def unwrapped(self):
        return self.sim.unwrapped if hasattr(self.sim, 'unwrapped') else self.sim


class SuperAgentView:
    """Per-step regrouping of the engine's [E, L] outputs into super agents; tensors live where the outputs live.

    groups: list of (id, [learner indices]); every learner that is not covered forms its own group."""

    def __init__(self, spec, super_agent_mapping, n_envs, device='cpu'):
        index = {aid: l for l, aid in enumerate(spec.learner_ids)}
        self.groups = [(sid, [index[m] for m in members]) for sid, members in super_agent_mapping.items()]
        covered = {l for _, ls in self.groups for l in ls}
        self.groups += [(spec.learner_ids[l], [l]) for l in range(spec.n_learners) if l not in covered]
        self.is_super = [sid in super_agent_mapping for sid, _ in self.groups]
        self.spec, self.E, self.L, self.Kg = spec, n_envs, spec.n_learners, len(self.groups)
        group_of = np.zeros(self.L, dtype=np.int64)
        for g, (_, ls) in enumerate(self.groups):
            group_of[ls] = g
        self.group_of = torch.from_numpy(group_of).to(device)
        self.uncovered = torch.from_numpy(np.array([l not in covered for l in range(self.L)])).to(device)   # no mask channel
        order = np.array([l for _, ls in self.groups for l in ls], dtype=np.int16)       # processing order of the inner sim
        self.order = torch.from_numpy(np.tile(order, (n_envs, 1))).to(device)
        self.seen_done = torch.zeros((n_envs, self.L), dtype=torch.bool, device=device)
        # null observation of every learner: -2 over its own window (observer.py:77-79,172-174,274-278), 0 in the row padding
        h, w, c, stride = spec.obs_shape()
        null = np.zeros((self.L, stride), dtype=np.int8)
        for l, a in enumerate(spec.learner_agents):
            if spec.klass[a] & K.AG_OBSERVING:
                n = 2 * int(spec.view_range[a]) + 1
                null[l, :(h * w if spec.observer == K.OBS_ABSOLUTE else n * n) * c] = -2
        self.null_rows = torch.from_numpy(null).to(device)

    def reset(self, env_mask=None):
        if env_mask is None:
            self.seen_done.zero_()
        else:
            self.seen_done[torch.as_tensor(env_mask, device=self.seen_done.device).bool()] = False

    def update(self, obs, reward, done, all_done):
        """obs [E, L, stride] int8, reward [E, L] f32, done [E, L] u8 (OUT_*), all_done [E] u8 (ENV_*) of the last step ->
        (obs with null rows, mask [E, L] bool, reward [E, G] f64, done [E, G] bool, valid [E, G] bool)."""
        fresh = (all_done & K.ENV_RESET) != 0                      # auto-reset happened in this call
        self.seen_done[fresh] = False
        valid = (done & K.OUT_VALID) != 0
        now_done = (done & K.OUT_DONE) != 0
        out_obs = obs.clone()
        if self.seen_done.any():
            out_obs = torch.where(self.seen_done[..., None], self.null_rows[None].expand_as(out_obs), out_obs)
        finished = self.seen_done | now_done
        mask = ~finished | self.uncovered
        E, G = obs.shape[0], self.Kg
        r = torch.zeros((E, G), dtype=torch.float64, device=obs.device)
        r.index_add_(1, self.group_of, torch.where(valid, reward, torch.zeros_like(reward)).double())
        members = torch.zeros((E, G), dtype=torch.int32, device=obs.device)
        members.index_add_(1, self.group_of, torch.ones((E, self.L), dtype=torch.int32, device=obs.device))
        n_fin = torch.zeros((E, G), dtype=torch.int32, device=obs.device)
        n_fin.index_add_(1, self.group_of, finished.int())
        n_seen = torch.zeros((E, G), dtype=torch.int32, device=obs.device)
        n_seen.index_add_(1, self.group_of, self.seen_done.int())
        group_done = n_fin == members
        group_valid = (n_seen < members) & ~fresh[:, None]          # the super agent was not done before this step
        self.seen_done |= now_done
        return out_obs, mask, r, group_done, group_valid

from abmarl_b200.sim.flatten import FlattenWrapper, FlattenActionWrapper, FlattenView   # noqa: E402,F401  (abmarl.sim.wrappers exports them too)
