"""PolicyNet: frame + frame_state + previous output + previous grid -> per-block logits
(reference policy/net.py:17-125).  Input features are built at 1/4 resolution (for 128-px blocks)
in fp32; trunk = ResNet-8 (width x2) followed by three stride-2 3x3 convs, so that one logit
comes out per block.  The net runs in TRAIN mode (batch statistics), see policy.py.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from blockcopy.core.frame import as_tensor
from blockcopy.policy.resnet import resnet8
from blockcopy.utils.profiler import timings


def build_policy_net_from_settings(settings: dict):
    return PolicyNet(block_size=settings["block_size"], task_num_classes=settings["block_num_classes"])


class PolicyNet(nn.Module):
    def __init__(self, block_size, task_num_classes) -> None:
        super().__init__()
        self.block_size = block_size
        self.scale_factor = 0.25 * 128 / self.block_size
        self.use_frame_state = True
        self.use_prev_output = True
        self.use_prev_grid = True
        self.task_num_classes = task_num_classes
        in_channels = 3 + (3 if self.use_frame_state else 0) + \
            (task_num_classes if self.use_prev_output else 0) + (1 if self.use_prev_grid else 0)
        self.backbone = resnet8(pretrained=False, in_channels=in_channels, width_factor=2)
        planes = 128
        self.layers = nn.Sequential(
            self._make_layer(self.backbone.OUT_CHANNELS, planes, relu=True),
            self._make_layer(planes, planes, relu=True),
            self._make_layer(planes, 1, relu=False),
        )
        # block_cuda_graphs: trunk forward and backward replayed as CUDA graphs (see _trunk_forward)
        self.use_cuda_graphs = False
        # NHWC activations / weights for the fp32 trunk (same math; cuDNN's NHWC TF32 kernels): set by
        # BlockCopyModel together with block_channels_last
        self.channels_last = False
        self.__dict__["_graphed"] = None  # (input shape, graphed callable); not a submodule: state_dict unchanged
        # frames that are not followed by a policy update need no autograd graph: their trunk forward runs on this
        # repo's kernels (policy/fused_net.py; fp16 operands, train-mode batch statistics).  False: always torch
        self.fused_inference = True
        self.__dict__["_fused"] = None
        # frames that ARE followed by a policy update: forward with saved activations + backward on this repo's kernels
        # (policy/fused_train.py) instead of torch autograd over cuDNN.  PolicyTrainRL sets it from
        # settings['block_policy_fused_training']
        self.fused_training = True
        self.__dict__["_trainer"] = None

    @staticmethod
    def _make_layer(cin, cout, kernel_size=3, stride=2, relu=True):
        mods = [nn.Conv2d(cin, cout, kernel_size=kernel_size, padding=(kernel_size - 1) // 2, stride=stride,
                          bias=not relu)]
        if relu:
            mods += [nn.BatchNorm2d(cout, momentum=0.02), nn.ReLU(inplace=False)]
        return nn.Sequential(*mods)

    def build_features(self, policy_meta: dict) -> torch.Tensor:
        """(N, 3+3+num_classes+1, H*s, W*s) fp32: nearest-resized frame | frame_state |
        previous output - 0.5 | previous grid - 0.5   (reference net.py:84-113)."""
        frame = as_tensor(policy_meta["inputs"])  # a lazily normalised U8Frame: the policy looks at the whole frame
        assert frame.dim() == 4 and frame.size(1) == 3
        fused = self._fused_features(policy_meta)
        if fused is not None:
            return fused
        feats = [F.interpolate(frame, scale_factor=self.scale_factor, mode="nearest").float()]
        size = feats[0].shape[2:]
        if self.use_frame_state:
            feats.append(F.interpolate(policy_meta["frame_state"], size=size, mode="nearest").float())
        if self.use_prev_output:
            assert policy_meta.get("output_repr", None) is not None
            rep = policy_meta["output_repr"]
            assert rep.dim() == 4
            feats.append(F.interpolate(rep, size=size, mode="nearest").type(feats[0].dtype) - 0.5)
        if self.use_prev_grid:
            assert policy_meta.get("grid", None) is not None
            g = policy_meta["grid"].type(feats[0].dtype)
            assert g.dim() == 4
            feats.append(F.interpolate(g, size=size, mode="nearest") - 0.5)
        return torch.cat(feats, dim=1).detach()

    def _fused_features(self, policy_meta: dict):
        """One sm_100a kernel (bc_policy_features) instead of 4 interpolates + casts + cat, when the
        inputs are the usual CUDA tensors; same values bit for bit (pure gathers and `- 0.5` in fp32)."""
        frame, state = as_tensor(policy_meta["inputs"]), policy_meta.get("frame_state", None)
        rep, grid = policy_meta.get("output_repr", None), policy_meta.get("grid", None)
        if not (self.use_frame_state and self.use_prev_output and self.use_prev_grid) or not frame.is_cuda:
            return None
        if state is None or rep is None or grid is None or rep.dim() != 4 or grid.dim() != 4:
            return None
        if frame.dtype not in (torch.float16, torch.float32) or state.dtype != frame.dtype or state.shape != frame.shape:
            return None
        if not frame.is_contiguous() or not state.is_contiguous() or rep.dtype not in (torch.float16, torch.float32):
            return None
        from blockcopy import _C

        return _C.policy_features(frame, state, rep, grid, self.scale_factor)

    def _trunk_forward(self, x: torch.Tensor) -> torch.Tensor:
        """backbone + head.  With ``use_cuda_graphs`` (BlockCopyModel sets it from ``block_cuda_graphs``) the
        ~45 fp32 launches of the forward and the ~90 of the backward are each one CUDA-graph replay
        (``torch.cuda.make_graphed_callables``: same kernels, same order, autograd-aware, so
        ``loss.backward()`` in PolicyTrainRL.optim works unchanged).  Capture runs the trunk a few times:
        the batch-norm running statistics are restored afterwards, so capturing is free of side effects."""
        if self.channels_last and x.is_cuda:
            x = x.contiguous(memory_format=torch.channels_last)
        if not (self.use_cuda_graphs and x.is_cuda and self.training and torch.is_grad_enabled()):
            return self.layers(self.backbone(x))
        g = self.__dict__["_graphed"]
        key = (tuple(x.shape), x.dtype, next(self.parameters()).data_ptr())
        if g is None or g[0] != key:
            trunk = nn.Sequential(self.backbone, self.layers)
            saved = {k: v.detach().clone() for k, v in trunk.state_dict().items()}
            grads = [None if q.grad is None else q.grad.detach().clone() for q in trunk.parameters()]
            fn = torch.cuda.make_graphed_callables(trunk, (x.detach().clone(),), allow_unused_input=True)
            with torch.no_grad():
                for k, v in trunk.state_dict().items():
                    v.copy_(saved[k])
                for q, gr in zip(trunk.parameters(), grads):
                    q.grad = gr
            g = self.__dict__["_graphed"] = (key, fn)
        return g[1](x)

    def _fused_trunk(self):
        t = self.__dict__["_fused"]
        if t is None:
            from blockcopy.policy.fused_net import FusedPolicyTrunk

            t = self.__dict__["_fused"] = FusedPolicyTrunk(self)
        return t

    def _fused_trainer(self):
        t = self.__dict__["_trainer"]
        if t is None:
            from blockcopy.policy.fused_train import FusedPolicyTrainer

            t = self.__dict__["_trainer"] = FusedPolicyTrainer(self)
        return t

    def _fused_forward(self, policy_meta: dict, train: bool = False):
        """Features written straight into the fused trunk's fp16 NHWC input plane (bc_policy_features_nhwc16) and
        the trunk behind it; None when the inputs are outside that kernel's envelope (see _fused_features)."""
        frame, state = as_tensor(policy_meta["inputs"]), policy_meta.get("frame_state", None)
        rep, grid = policy_meta.get("output_repr", None), policy_meta.get("grid", None)
        if not (self.use_frame_state and self.use_prev_output and self.use_prev_grid) or not frame.is_cuda:
            return None
        if state is None or rep is None or grid is None or rep.dim() != 4 or grid.dim() != 4 or frame.dim() != 4:
            return None
        if frame.dtype not in (torch.float16, torch.float32) or state.dtype != frame.dtype or state.shape != frame.shape:
            return None
        if not frame.is_contiguous() or not state.is_contiguous() or rep.dtype not in (torch.float16, torch.float32):
            return None
        from blockcopy import _C

        fused = self._fused_trainer() if train else self._fused_trunk()
        shape = _C.policy_features_shape(frame, rep, self.scale_factor)
        if not fused.supports_shape(shape, frame.device):
            return None
        if train:
            fused.use_cuda_graph = self.use_cuda_graphs
            with timings.env("policy/net/layers", 5):
                return fused.run_train(lambda x16: _C.policy_features_nhwc16(x16, frame, state, rep, grid, self.scale_factor),
                                       shape, frame.device)
        with timings.env("policy/net/layers", 5):
            return fused.run(lambda x16: _C.policy_features_nhwc16(x16, frame, state, rep, grid, self.scale_factor),
                             shape, frame.device, use_cuda_graph=self.use_cuda_graphs)

    def forward(self, policy_meta: dict, no_grad: bool = False):
        """no_grad: the caller will not back-propagate through this call (a frame without policy update): the
        trunk may run on the inference kernels."""
        N, C, H, W = policy_meta["inputs"].shape
        logits = None
        if no_grad and self.fused_inference and self.training:
            logits = self._fused_forward(policy_meta)
        elif not no_grad and self.fused_training and self.training and torch.is_grad_enabled():
            logits = self._fused_forward(policy_meta, train=True)
        if logits is None:
            with timings.env("policy/net/build_features", 5):
                x = self.build_features(policy_meta)
            with timings.env("policy/net/layers", 5):
                fused = self._fused_trunk() if (no_grad and self.fused_inference and self.training and x.is_cuda) else None
                if fused is not None and fused.supports(x):
                    logits = fused(x, use_cuda_graph=self.use_cuda_graphs)
                else:
                    logits = self._trunk_forward(x)
        expect = (N, 1, H // self.block_size, W // self.block_size)
        assert logits.shape == expect, f"logits shape: {logits.shape}, frame shape: {(N, C, H, W)}, " \
                                       f"block size: {self.block_size}"
        return logits
