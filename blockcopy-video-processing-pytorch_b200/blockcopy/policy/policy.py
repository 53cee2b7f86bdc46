"""Block-execution policies (reference policy/policy.py:14-370).

A policy maps ``policy_meta`` (inputs, previous outputs, frame state, previous grid) to a boolean
execution grid ``(N,1,GH,GW)``; ``Policy`` is the plug-in interface -- consumers either build one
from the settings dict or assign ``model.policy = MyPolicy(...)``.  ``forward`` must set ``grid``
and call ``self.stats.add_policy_meta``; ``optim`` is called once per frame after the model ran.

Numerics follow SURVEY.md A.5: Bernoulli sampling of the logits, executed-block count rounded UP
to a multiple of ``int(G/16)`` with Python's ``random`` module (so that ``random.seed`` reproduces
the reference's masks), REINFORCE with reward = information gain + gamma * s|s|,
s = -(running_cost - target), RMSprop.
"""
import abc
import logging
import random
from abc import abstractmethod

import torch
import torch.nn.functional as F
from torch.distributions import Bernoulli

from blockcopy.policy.information_gain import InformationGain, InformationGainObjectDetection, InformationGainSemSeg
from blockcopy.policy.net import PolicyNet, build_policy_net_from_settings
from blockcopy.utils.hints import get_num_exec_hint, set_num_exec_hint
from blockcopy.utils.profiler import timings


def build_policy_from_settings(settings: dict):
    """Policy object for ``settings['block_policy']`` in {all, none, random, rl_semseg, rl_objectdetection}."""
    name = settings["block_policy"]
    logging.info(f"> Policy: {name} with execution percentage target {settings['block_target']} "
                 f"and block size {settings['block_size']}")
    quantize = 1 / 16
    common = dict(block_size=settings["block_size"], verbose=settings["block_policy_verbose"])
    if name == "all":
        return PolicyAll(**common)
    if name == "none":
        return PolicyNone(**common)
    if name == "random":
        return PolicyRandom(quantize_number_exec=quantize, **common)
    if name.startswith("rl_"):
        net = build_policy_net_from_settings(settings)
        net.fused_inference = bool(settings.get("block_policy_fused", True))  # policy/fused_net.py (not in the reference)
        net.fused_training = bool(settings.get("block_policy_fused_training", True))  # policy/fused_train.py
        optimizer = build_policy_optimizer_from_settings(settings, net)
        if name == "rl_semseg":
            ig = InformationGainSemSeg(num_classes=settings["block_num_classes"])
        elif name == "rl_objectdetection":
            ig = InformationGainObjectDetection(num_classes=settings["block_num_classes"])
        else:
            raise AttributeError(f'Policy with name "{name}" not defined!')
        return PolicyTrainRL(block_target=settings["block_target"], cost_momentum=settings["block_cost_momentum"],
                             optimizer=optimizer, complexity_weight=settings["block_complexity_weight"],
                             quantize_number_exec=quantize, policy_net=net, information_gain=ig,
                             shared_across_ranks=bool(settings.get("block_policy_shared", False)),
                             device_sampling=bool(settings.get("block_policy_device_sampling", True)),
                             strict_checks=bool(settings.get("block_policy_strict_checks", False)), **common)
    raise NotImplementedError(f"Policy {name} not implemented")


def build_policy_optimizer_from_settings(settings: dict, net: PolicyNet) -> torch.optim.Optimizer:
    # torch.optim.RMSprop with a one-kernel step (policy/fused_optim.py): same interface, hyper-parameters and state
    from blockcopy.policy.fused_optim import FusedRMSprop

    return FusedRMSprop(net.parameters(), lr=settings["block_optim_lr"],
                        weight_decay=settings["block_optim_wd"], centered=False,
                        momentum=settings["block_optim_momentum"])


def sample_grid_host(probs, uniforms, multiple: int, at_least_one: bool = False):
    """Host restatement (numpy) of bc_sample_grid: probs (G,), uniforms (2G,) float32 -> (grid bool (G,), executed
    after rounding, executed before).  Bernoulli draw ``u < p``; the executed count is rounded up to
    ``multiple * (1 + (E0 - 1) // multiple)`` exactly like the reference (policy.py:139-140) and the extra cells are
    the skipped ones with the smallest keys ``uniforms[G + g]`` (ties: lower index) -- a uniform random subset, where
    the reference uses ``random.sample`` (policy.py:141)."""
    import numpy as np

    p = np.asarray(probs, dtype=np.float32).reshape(-1)
    G = p.size
    u = np.asarray(uniforms, dtype=np.float32).reshape(-1)
    grid = u[:G] < p
    if at_least_one and not grid.any():
        grid[0] = True
    e0 = int(grid.sum())
    target = e0
    if multiple > 0:
        target = multiple * (1 + (e0 - 1) // multiple)  # Python floor division: e0 == 0 gives 0
    need = min(target - e0, G - e0)
    if need > 0:
        skipped = np.nonzero(~grid)[0]
        order = np.lexsort((skipped, u[G:2 * G][skipped]))  # by key, then by index
        grid[skipped[order[:need]]] = True
    return grid, e0 + max(need, 0), e0


class PolicyStats:
    """Running executed-block fraction; also fills num_exec / num_total / perc_exec of policy_meta."""

    def __init__(self):
        self.count_images = 0
        self.exec = 0
        self.total = 0

    def add_policy_meta(self, policy_meta: dict) -> dict:
        grid = policy_meta["grid"]
        num_exec = get_num_exec_hint(grid)             # host-generated grids carry their (still valid) count
        if num_exec is None:
            num_exec = int(grid.sum())                 # the one host sync per frame
            set_num_exec_hint(grid, num_exec)          # reused by TensorWrapper.to_blocks while the grid is unchanged
        num_total = int(grid.numel())
        policy_meta["num_exec"] = num_exec
        policy_meta["num_total"] = num_total
        policy_meta["perc_exec"] = float(num_exec) / num_total
        self.count_images += grid.size(0)
        self.exec += num_exec
        self.total += num_total
        return policy_meta

    def get_exec_percentage(self):
        return float(self.exec) / self.total

    def __repr__(self) -> str:
        return f"Policy stats: average exec percentage [0 - 1] : {self.get_exec_percentage():0.3f}"


class Policy(torch.nn.Module, metaclass=abc.ABCMeta):
    """Plug-in interface for execution policies."""

    def __init__(self, block_size, verbose=False, quantize_number_exec=0):
        super().__init__()
        self.block_size = block_size
        self.net = None
        self.optimizer = None
        self.verbose = verbose
        self.stats = PolicyStats()
        self.fp16_enabled = False
        self.quantize_number_exec = quantize_number_exec

    def is_trainable(self):
        return self.net is not None

    def _grid_shape(self, policy_meta: dict):
        N, C, H, W = policy_meta["inputs"].shape
        assert H % self.block_size == 0, f"input height ({H}) not a multiple of block size {self.block_size}!"
        assert W % self.block_size == 0, f"input width  ({W}) not a multiple of block size {self.block_size}!"
        return (N, 1, H // self.block_size, W // self.block_size)

    def quantize_number_exec_grid(self, grid: torch.Tensor) -> torch.Tensor:
        """Round the number of executed blocks UP to a multiple of ``quantize_number_exec * G`` by
        switching on randomly chosen skipped blocks (bounded set of tile-batch shapes => bounded
        set of CUDA graphs / cuDNN plans).  Uses ``random.sample`` like the reference
        (policy.py:124-144), so seeding ``random`` reproduces its choice."""
        if self.quantize_number_exec > 0:
            with timings.env("policy/quantize_number_exec", 3):
                host = grid.detach().reshape(-1).bool().cpu().numpy()  # the one device round trip (G bytes)
                skipped = (~host).nonzero()[0].tolist()                 # ascending, like torch.nonzero
                total = host.size
                num_exec = total - len(skipped)
                multiple = int(total * self.quantize_number_exec)
                target = multiple * (1 + (num_exec - 1) // multiple)
                extra = random.sample(skipped, target - num_exec)
                if extra:
                    host[extra] = True  # one upload of the whole (G-byte) mask instead of index upload + index_put
                    grid.copy_(torch.from_numpy(host).view(grid.shape), non_blocking=False)
                set_num_exec_hint(grid, target)  # counted on the host just now: saves the device round trip
        return grid

    @abstractmethod
    def forward(self, policy_meta: dict) -> dict:
        raise NotImplementedError

    def optim(self, policy_meta, train=True, **kwargs):
        return policy_meta


class PolicyAll(Policy):
    """Execute every block of every frame."""

    def forward(self, policy_meta: dict) -> dict:
        shape = self._grid_shape(policy_meta)
        grid = torch.ones(shape, device=policy_meta["inputs"].device, dtype=torch.bool)
        set_num_exec_hint(grid, grid.numel())  # known on the host: no device round trip
        policy_meta["grid"] = grid
        return self.stats.add_policy_meta(policy_meta)


class PolicyNone(Policy):
    """Execute nothing once a previous output exists (the first TWO frames run fully: the test is
    on ``outputs_prev``, reference policy.py:189)."""

    def forward(self, policy_meta: dict) -> dict:
        shape = self._grid_shape(policy_meta)
        first = policy_meta.get("outputs_prev", None) is None
        grid = torch.full(shape, bool(first), device=policy_meta["inputs"].device, dtype=torch.bool)
        set_num_exec_hint(grid, grid.numel() if first else 0)
        policy_meta["grid"] = grid
        return self.stats.add_policy_meta(policy_meta)


class PolicyRandom(Policy):
    """Each block executes with probability 1/2 (then quantised); first two frames run fully."""

    def forward(self, policy_meta: dict) -> dict:
        shape = self._grid_shape(policy_meta)
        dev = policy_meta["inputs"].device
        if policy_meta.get("outputs_prev", None) is None:
            grid = torch.ones(shape, device=dev).type(torch.bool)
        else:
            grid = (torch.randn(shape, device=dev) > 0).type(torch.bool)
        policy_meta["grid"] = self.quantize_number_exec_grid(grid)
        return self.stats.add_policy_meta(policy_meta)


class PolicyTrainRL(Policy, metaclass=abc.ABCMeta):
    """REINFORCE policy trained online at test time (reference policy.py:219-370)."""

    def __init__(self, block_size: int, block_target: float, optimizer: torch.optim.Optimizer,
                 complexity_weight: float, policy_net: PolicyNet, information_gain: InformationGain,
                 cost_momentum: float = 0.9, at_least_one: bool = False, quantize_number_exec: float = 0,
                 verbose: bool = False, shared_across_ranks: bool = False, device_sampling: bool = True,
                 strict_checks: bool = False):
        super().__init__(block_size, verbose, quantize_number_exec)
        # Bernoulli draw + count quantisation in one kernel (bc_sample_grid), the only host round trip of the frame
        # being the executed-block count the API exposes anyway.  False: the reference's host procedure
        # (grid download, Python random.sample, mask upload) -- reproduces its masks under random.seed.
        self.device_sampling = device_sampling
        # One policy shared by the streams of all ranks (settings key block_policy_shared, not in the reference):
        # the only collective of the whole path -- a sum of the flat policy-gradient buffer (2.4 MB fp32) every
        # block_train_interval frames; with equal initial weights (broadcast from rank 0) and equal gradients the
        # replicas stay identical by construction.  Every rank must train on the same frames.
        self.shared_across_ranks = shared_across_ranks
        self._shared_synced = False
        assert 0 <= block_target <= 1
        self.block_target = block_target
        self.information_gain = information_gain
        self.momentum = cost_momentum
        self.running_cost = None
        self.net = policy_net
        self.complexity_weight_gamma = complexity_weight
        self.optimizer = optimizer
        self.at_least_one = at_least_one
        self.strict_checks = strict_checks
        self._deferred = []

    def forward(self, policy_meta: dict):
        shape = self._grid_shape(policy_meta)
        if policy_meta["outputs"] is None:
            # no temporal history yet: execute everything
            grid = torch.ones(shape, device=policy_meta["inputs"].device, dtype=torch.bool)
            set_num_exec_hint(grid, grid.numel())
            policy_meta["grid"] = grid
        else:
            if self.shared_across_ranks and not self._shared_synced:
                self.sync_shared_policy()
            with torch.enable_grad():
                with timings.env("policy/net", 3):
                    assert self.net.training
                    # BlockCopyModel announces whether optim() will train on this frame; if not, nothing is
                    # back-propagated through this forward (grid_log_probs is only read by the training step)
                    no_grad = policy_meta.get("policy_will_train", True) is False and \
                        getattr(self.net, "fused_inference", False)
                    grid_logits = self.net(policy_meta, no_grad=no_grad) if no_grad else self.net(policy_meta)
                    self._check(torch.isnan(grid_logits).any(), "Policy net returned NaN's, maybe optimization problem?")
                sampled = None
                if self.device_sampling and grid_logits.is_cuda and grid_logits.numel() <= 8192:
                    with timings.env("policy/sample", 3):
                        sampled = self._sample_on_device(grid_logits, no_grad)
                if sampled is not None:
                    grid, probs, dist = sampled
                    policy_meta["grid_log_probs"] = dist.log_prob(grid.to(grid_logits.dtype)) if dist is not None else None
                    policy_meta["grid_probs"] = probs
                    policy_meta["grid"] = grid
                    return self.stats.add_policy_meta(policy_meta)
                with timings.env("policy/sample", 3):
                    if no_grad:
                        # same draw as Bernoulli(logits=...).sample() (= torch.bernoulli(sigmoid(logits))) without
                        # building the distribution object; log-probabilities are only read by a training step
                        dist, probs = None, torch.sigmoid(grid_logits)
                        grid = torch.bernoulli(probs)
                    else:
                        dist = Bernoulli(logits=grid_logits)
                        grid = dist.sample()
                        probs = dist.probs
                if self.at_least_one and grid.sum() == 0:
                    grid[0, 0, 0, 0] = 1
                self.flush_checks()  # this path talks to the host anyway
                grid = self.quantize_number_exec_grid(grid)
                policy_meta["grid_log_probs"] = dist.log_prob(grid) if dist is not None else None
                policy_meta["grid_probs"] = probs
                assert grid.dim() == 4 and probs.shape == grid.shape
                hint = get_num_exec_hint(grid)
                grid = grid.bool()
                if hint is not None:
                    set_num_exec_hint(grid, hint)
                policy_meta["grid"] = grid
        return self.stats.add_policy_meta(policy_meta)

    def _sample_on_device(self, grid_logits: torch.Tensor, no_grad: bool):
        """(bool grid with its count hint, probabilities, Bernoulli distribution | None)."""
        from blockcopy import _C

        if no_grad:
            dist, probs = None, torch.sigmoid(grid_logits.detach())
        else:
            # argument / sample validation of torch.distributions is a host sync each (torch._is_all_true); NaN logits are
            # caught by the (deferred) check in forward()
            dist = Bernoulli(logits=grid_logits, validate_args=self.strict_checks)
            probs = dist.probs
        p32 = probs.detach().float().contiguous()
        G = p32.numel()
        uniforms = torch.rand(2 * G, device=p32.device, dtype=torch.float32)  # torch's generator: torch.manual_seed
        multiple = int(G * self.quantize_number_exec) if self.quantize_number_exec > 0 else 0
        grid, counts = _C.sample_grid(p32, uniforms, multiple, self.at_least_one)
        # the one host sync of the frame (the API's num_exec is a Python int); checks deferred since the last one
        # (NaN asserts, the "not well trained" diagnostic) ride on the same device -> host read
        set_num_exec_hint(grid, self._read_count_and_checks(counts))
        return grid, probs, dist

    # ------------------------------------------------------------------ device-side checks without extra host syncs
    def _check(self, bad: torch.Tensor, message: str, warn: bool = False):
        """`bad` (0-d bool tensor) must be False.  On the GPU the verdict is not read here -- that would stall the host
        until the frame's kernels have run, once per check -- but together with the next executed-block count (the one
        device -> host read every rl frame needs anyway): an AssertionError (the reference asserts in place,
        policy.py:253,335,345) or, with `warn`, a printed warning arrives at most one frame later.
        ``strict_checks`` (settings['block_policy_strict_checks']) restores the immediate behaviour."""
        if not bad.is_cuda or self.strict_checks:
            if bool(bad):
                if not warn:
                    raise AssertionError(message)
                print(message)
            return
        self._deferred.append((bad.detach().reshape(1).to(torch.int32), message, warn))
        if len(self._deferred) > 32:  # nobody is reading counts (host-side sampling path): do not pile up
            self.flush_checks()

    def _read_count_and_checks(self, counts: torch.Tensor) -> int:
        """int(counts[0]) plus every deferred check, in one device -> host copy."""
        if not self._deferred:
            return int(counts[0])
        pending, self._deferred = self._deferred, []
        vals = torch.cat([counts[:1]] + [f for f, _, _ in pending]).tolist()
        for v, (_, message, warn) in zip(vals[1:], pending):
            if v:
                if not warn:
                    raise AssertionError(message)
                print(message)
        return int(vals[0])

    def flush_checks(self):
        """Read the deferred checks now (end of a clip / before saving a checkpoint)."""
        if self._deferred:
            self._read_count_and_checks(torch.zeros(1, dtype=torch.int32, device=self._deferred[0][0].device))

    def _shared_world(self):
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2:
            return None
        return dist

    def sync_shared_policy(self):
        """Shared policy: every rank starts from rank 0's weights and buffers (called lazily before the first
        forward; call it explicitly after loading a checkpoint on rank 0)."""
        dist = self._shared_world()
        self._shared_synced = True
        if dist is None:
            return
        with torch.no_grad():
            for t in list(self.net.parameters()) + list(self.net.buffers()):
                dist.broadcast(t.data, src=0)
        from blockcopy.policy.fused_optim import PARAM_EPOCH

        PARAM_EPOCH[0] += 1  # written through .data: the fused trunks must re-pack their fp16 copies

    def _allreduce_gradients(self):
        """Average of the policy gradients over all ranks, one flat buffer, in place."""
        dist = self._shared_world()
        if dist is None:
            return
        grads = [q.grad for q in self.net.parameters() if q.grad is not None]
        if not grads:
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(dist.get_world_size())
        off = 0
        for g in grads:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n

    def _get_information_gain(self, policy_meta: dict) -> torch.Tensor:
        with timings.env("policy/information_gain", 3):
            ig = self.information_gain(policy_meta)
            assert ig.dim() == 4
            return ig

    def _get_reward_complexity(self, policy_meta: dict) -> float:
        s = -float(self.running_cost - self.block_target)
        return s * abs(s)

    def optim(self, policy_meta: dict, train=True) -> dict:
        policy_meta["output_repr"] = self.information_gain.get_output_repr(policy_meta)
        grid = policy_meta["grid"]
        assert grid.dim() == 4
        block_use = policy_meta["perc_exec"]
        if self.running_cost is None:
            self.running_cost = block_use
        self.running_cost = self.running_cost * self.momentum + (1 - self.momentum) * block_use

        if policy_meta["outputs_prev"] is not None and train:
            with torch.enable_grad():
                ig = self._get_information_gain(policy_meta)
                policy_meta["information_gain"] = ig
                reward_complexity_weighted = self._get_reward_complexity(policy_meta) * self.complexity_weight_gamma
                reward = ig + reward_complexity_weighted
                assert reward.dim() == 4
                self._check(torch.isnan(reward).any(), "PolicyTrainRL.optim: NaN in the reward")
                log_probs = policy_meta["grid_log_probs"]
                if log_probs is None:
                    raise RuntimeError(
                        "PolicyTrainRL.optim(train=True) on a frame whose forward ran without autograd: "
                        "policy_meta['policy_will_train'] was False for this frame (BlockCopyModel sets it from "
                        "clip_length % block_train_interval; a custom driver must set it to True, or leave it out, "
                        "on the frames it trains on)")
                reward = F.adaptive_max_pool2d(reward, output_size=log_probs.shape[2:])
                reward = torch.where(grid, reward, -reward)  # skipped blocks: negated reward
                loss_policy = (-log_probs * reward.detach()).mean()
                self._check(torch.isnan(loss_policy), "PolicyTrainRL.optim: the policy loss is NaN")
                with timings.env("policy/optimizer_backward", 3):
                    loss_policy.backward()
                if self.shared_across_ranks:
                    with timings.env("policy/allreduce", 3):
                        self._allreduce_gradients()
                with timings.env("policy/optimizer_step", 3):
                    self.optimizer.step()
                    self.optimizer.zero_grad(set_to_none=True)

                if self.stats.count_images > 300 and not self.verbose and not self.strict_checks \
                        and policy_meta["grid_probs"].is_cuda:
                    # the reference's diagnostic without its two host round trips (boolean-mask indexing + the comparison):
                    # masked means on the device, the verdict read with the next frame's executed-block count
                    probs, g = policy_meta["grid_probs"].detach().float(), grid.to(torch.float32)
                    n_exec = g.sum()
                    exec_mean = (probs * g).sum() / n_exec
                    skip_mean = (probs * (1 - g)).sum() / (g.numel() - n_exec)
                    self._check(exec_mean - skip_mean < 0.3, "Warning: Block execution policy seems not well trained yet.",
                                warn=True)
                elif self.verbose or self.stats.count_images > 300:
                    exec_mean = policy_meta["grid_probs"][grid].mean()
                    skip_mean = policy_meta["grid_probs"][~grid].mean()
                    if self.verbose:
                        print(f"BLOCKS/running_cost: {self.running_cost: 0.3f} \n"
                              f"BLOCKS/block_use: {block_use:0.3f} \n"
                              f"BLOCKS/information_gain_max: {ig.max()} \n"
                              f"BLOCKS/information_gain_min: {ig.min()} \n"
                              f"BLOCKS/reward_complexity_weighted: {reward_complexity_weighted} \n"
                              f"BLOCKS/avg_prob_exec: {exec_mean:0.3f} \n"
                              f"BLOCKS/avg_prob_skip: {skip_mean:0.3f} \n")
                        print(self.stats)
                    if self.stats.count_images > 300 and exec_mean - skip_mean < 0.3:
                        print("Warning: Block execution policy seems not well trained yet.")
        return policy_meta
