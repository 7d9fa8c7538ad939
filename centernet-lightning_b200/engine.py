"""Host wrapper of the sm_100a forward engine (csrc/cnl_conv.cu) - the replacement for
``GenericModel.forward`` (reference models/meta.py:41-47).

An ``Engine`` owns: the C-side plan (``cnl_engine``), one arena of device memory holding packed
weights + every activation buffer, and optionally a CUDA graph of the whole forward (+ decode).
torch provides the memory, the stream and graph capture; all arithmetic is in libcnl_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib
from .plan import Plan

PRECISION_SPLIT = 0   # fp16 hi+lo operands, 3 tensor-core passes, fp32 accumulate: fp32-equivalent (parity default)
_KINDS = {"conv": 0, "stem": 1, "dw": 2, "fuse": 3, "stem3x3": 4}     # cnl_conv_desc.kind (include/cnl_b200.h)
PRECISION_FAST = 1    # single fp16 pass: ~1e-1 logit error on a 50-conv network, reported separately


class Engine:
    def __init__(self, plan: Plan, batch: int, height: int, width: int, device: torch.device,
                 precision: int = PRECISION_SPLIT):
        if device.type != "cuda":
            raise RuntimeError("the cnl_b200 engine runs on CUDA (sm_100a) only - there is no CPU fallback")
        self.lib = _lib.load()
        self.plan = plan
        self.batch, self.height, self.width = int(batch), int(height), int(width)
        self.device = device
        self.precision = int(precision)
        self.buffer_ids: Dict[str, int] = {name: i for i, name in enumerate(plan.buffers)}
        if list(plan.buffers)[0] != "image":
            raise ValueError("plan buffer 0 must be the image")
        bufs = (_lib.BufferDesc * len(plan.buffers))()
        for i, b in enumerate(plan.buffers.values()):
            bufs[i] = _lib.BufferDesc(b.channels, b.stride, int(b.fp32_nchw))
        ops = (_lib.ConvDesc * len(plan.ops))()
        self._keep: List[np.ndarray] = []
        for i, op in enumerate(plan.ops):
            w = np.ascontiguousarray(op.weight.detach().cpu().numpy(), dtype=np.float32)
            b = np.ascontiguousarray(op.bias.detach().cpu().numpy(), dtype=np.float32)
            self._keep += [w, b]
            srcs = [self.buffer_ids[s] for s in (op.srcs or [])] + [-1, -1, -1]
            scales = list(op.scales or []) + [0.0, 0.0, 0.0]
            ops[i] = _lib.ConvDesc(
                _KINDS[op.kind], self.buffer_ids[op.src], self.buffer_ids[op.dst], op.cin, op.cout,
                op.ksize, op.stride, op.pad, int(op.relu), op.src_c_off, op.dst_c_off,
                self.buffer_ids[op.residual] if op.residual is not None else -1, op.residual_up,
                op.kh, op.kw, op.pad_h, op.pad_w, op.dst_up, op.dst_phase,
                w.ctypes.data, b.ctypes.data, srcs[1], srcs[2], scales[0], scales[1], scales[2], op.resize)
        handle = C.c_void_p()
        dev_index = device.index if device.index is not None else torch.cuda.current_device()
        st = self.lib.cnl_engine_create(C.byref(handle), bufs, len(plan.buffers), ops, len(plan.ops),
                                        self.batch, self.height, self.width, self.precision, dev_index)
        _lib.check(st, "cnl_engine_create")
        self.handle = handle
        self._keep.clear()                                   # weights were packed inside create
        nbytes = self.lib.cnl_engine_arena_bytes(self.handle)
        with torch.cuda.device(device):
            raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
            shift = (-raw.data_ptr()) % 1024
            self._raw = raw
            self.arena = raw[shift:shift + nbytes]
            stream = torch.cuda.current_stream(device)
            st = self.lib.cnl_engine_upload(self.handle, self.arena.data_ptr(), stream.cuda_stream)
        _lib.check(st, "cnl_engine_upload")
        self.outputs: Dict[str, torch.Tensor] = {}
        for head, bname in plan.outputs.items():
            self.outputs[head] = self._view_fp32(bname)
        self.num_ops = len(plan.ops)
        self.last_launches = 0

    # ------------------------------------------------------------------------------------------
    def _view_fp32(self, bname: str) -> torch.Tensor:
        b = self.plan.buffers[bname]
        off = self.lib.cnl_engine_buffer_offset(self.handle, self.buffer_ids[bname])
        h, w = self.height // b.stride, self.width // b.stride
        n = self.batch * b.channels * h * w
        return self.arena[off:off + 4 * n].view(torch.float32).view(self.batch, b.channels, h, w)

    def forward(self, image: torch.Tensor, first_op: int = 0, last_op: Optional[int] = None,
                sigmoid_head: Optional[str] = None) -> Dict[str, torch.Tensor]:
        """Runs the op range on the current stream.  Returns views of the head outputs inside the arena
        (valid until the next forward).  ``sigmoid_head``: that head's out-conv stores probabilities (the reference's
        ``.sigmoid()`` of models/centernet.py:205 fused into its epilogue) instead of logits."""
        if image is not None:
            if tuple(image.shape) != (self.batch, 3, self.height, self.width):
                raise ValueError(f"image must be {(self.batch, 3, self.height, self.width)}, got {tuple(image.shape)}")
            if image.dtype != torch.float32 or not image.is_cuda or not image.is_contiguous():
                raise ValueError("image must be a contiguous float32 CUDA tensor (NCHW)")
        if self.handle is None:
            raise RuntimeError("this engine was closed (its model's weights were reloaded): ask the model for a new one")
        n = C.c_int(0)
        with torch.cuda.device(self.device):        # launches go to the engine's device whatever the caller's current one is
            sig = self.buffer_ids[self.plan.outputs[sigmoid_head]] if sigmoid_head is not None else -1
            st = self.lib.cnl_engine_forward_act(self.handle, self.arena.data_ptr(), image.data_ptr() if image is not None else None,
                                                 first_op, self.num_ops if last_op is None else last_op, sig,
                                                 torch.cuda.current_stream(self.device).cuda_stream, C.byref(n))
        _lib.check(st, "cnl_engine_forward")
        self.last_launches = n.value
        return self.outputs

    def kernel_forms(self) -> Dict[str, str]:
        """op name -> "rows" (row-rolling A operand) / "pair" (cta_group::2) / "tile" (single-CTA im2col tiles)."""
        out = {}
        for i, op in enumerate(self.plan.ops):
            f = self.lib.cnl_engine_op_form(self.handle, i)
            out[op.name] = "rows" if f & 1 else ("pair" if f & 2 else "tile")
        return out

    def read_buffer(self, name: str) -> torch.Tensor:
        """(N,C,H,W) fp32 copy of any activation buffer (hi+lo planes summed) - for tests."""
        b = self.plan.buffers[name]
        h, w = self.height // b.stride, self.width // b.stride
        out = torch.empty((self.batch, b.channels, h, w), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            st = self.lib.cnl_engine_read_buffer(self.handle, self.arena.data_ptr(), self.buffer_ids[name], out.data_ptr(),
                                                 torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(st, "cnl_engine_read_buffer")
        return out

    def write_buffer(self, name: str, value: torch.Tensor) -> None:
        b = self.plan.buffers[name]
        h, w = self.height // b.stride, self.width // b.stride
        value = value.to(device=self.device, dtype=torch.float32).contiguous()
        if tuple(value.shape) != (self.batch, b.channels, h, w):
            raise ValueError(f"{name}: expected {(self.batch, b.channels, h, w)}, got {tuple(value.shape)}")
        with torch.cuda.device(self.device):
            st = self.lib.cnl_engine_write_buffer(self.handle, self.arena.data_ptr(), self.buffer_ids[name], value.data_ptr(),
                                                  torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(st, "cnl_engine_write_buffer")

    def close(self) -> None:
        if getattr(self, "handle", None) is not None:
            self.lib.cnl_engine_destroy(self.handle)
            self.handle = None
            self.arena = self._raw = None             # release the activation arena with the plan

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
