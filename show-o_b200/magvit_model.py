"""`MAGVITv2`: drop-in for the reference's models.MAGVITv2 (models/modeling_magvitv2.py:402-433) on libshowo_b200.so.

`get_code(pixel_values) -> LongTensor[B, N]` and `decode_code(ids, shape=None) -> FloatTensor[B,3,R,R]` keep the
reference's signatures; parameters are held as plain fp32 tensors under the reference's state_dict names and streamed
into the engine (NHWC bf16 implicit-GEMM convolutions on tcgen05, see csrc/magvit.cu).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib


def _param_shapes() -> Dict[str, tuple]:
    """Every parameter of MAGVITv2() in the reference's naming (encoder :59-141, decoder :278-362)."""
    ch, z = 128, 13
    S: Dict[str, tuple] = {}

    def conv(p, o, i, k):
        S[p + ".weight"] = (o, i, k, k)
        S[p + ".bias"] = (o,)

    def norm(p, c):
        S[p + ".weight"] = (c,)
        S[p + ".bias"] = (c,)

    def res(p, i, o):
        norm(p + ".norm1", i); conv(p + ".conv1", o, i, 3); norm(p + ".norm2", o); conv(p + ".conv2", o, o, 3)
        if i != o:
            conv(p + ".nin_shortcut", o, i, 1)

    def attn(p, c):
        norm(p + ".norm", c)
        for n in ("q", "k", "v", "proj_out"):
            conv(p + "." + n, c, c, 1)

    enc_mult, enc_blocks = [1, 2, 2, 4, 4], [4, 3, 4, 3, 4]
    dec_mult, dec_blocks = [1, 1, 2, 2, 4], [4, 4, 3, 4, 3]
    conv("encoder.conv_in", ch, 3, 3)
    bi = ch
    for lvl in range(5):
        bo = ch * enc_mult[lvl]
        for blk in range(enc_blocks[lvl]):
            res(f"encoder.down.{lvl}.block.{blk}", bi, bo)
            bi = bo
        if lvl != 4:
            conv(f"encoder.down.{lvl}.downsample.conv", bi, bi, 3)
    res("encoder.mid.block_1", bi, bi); attn("encoder.mid.attn_1", bi); res("encoder.mid.block_2", bi, bi)
    norm("encoder.norm_out", bi); conv("encoder.conv_out", z, bi, 3); conv("encoder.quant_conv", z, z, 1)
    bi = ch * dec_mult[4]
    conv("decoder.conv_in", bi, z, 3)
    res("decoder.mid.block_1", bi, bi); attn("decoder.mid.attn_1", bi); res("decoder.mid.block_2", bi, bi)
    for lvl in reversed(range(5)):
        bo = ch * dec_mult[lvl]
        for blk in range(dec_blocks[lvl]):
            res(f"decoder.up.{lvl}.block.{blk}", bi, bo)
            bi = bo
        if lvl != 0:
            conv(f"decoder.up.{lvl}.upsample.conv", bi, bi, 3)
    norm("decoder.norm_out", bi); conv("decoder.conv_out", 3, bi, 3); conv("decoder.post_quant_conv", z, z, 1)
    return S


class MAGVITv2(nn.Module):
    def __init__(self, materialize: bool = True):
        super().__init__()
        self._names = {}
        if materialize:
            for name, shp in _param_shapes().items():
                t = torch.empty(shp)
                if len(shp) == 4:
                    nn.init.kaiming_uniform_(t, a=math.sqrt(5))
                elif ".norm" in name:
                    t.fill_(1.0 if name.endswith("weight") else 0.0)
                else:
                    t.zero_()
                key = name.replace(".", "__")
                self._names[key] = name
                self.register_parameter(key, nn.Parameter(t, requires_grad=False))
        # LFQuantizer buffers (modeling_magvitv2.py:186-197) for callers that read them
        idx = torch.arange(8192)
        bits = (idx.unsqueeze(1) >> torch.arange(12, -1, -1)) & 1
        self.register_buffer("quantize_embedding", bits.float() * 2 - 1, persistent=False)
        self._engine = None
        self._versions = None
        self._streamed = False

    # state_dict under the reference's dotted names (nn.Module forbids dots in parameter names: they are registered with "__")
    def state_dict(self, *args, destination=None, prefix="", keep_vars=False, **kwargs):
        if args:                         # legacy positional form (destination, prefix, keep_vars)
            destination = args[0]
            prefix = args[1] if len(args) > 1 else prefix
            keep_vars = args[2] if len(args) > 2 else keep_vars
        out = destination if destination is not None else {}
        for k, v in self._parameters.items():
            out[prefix + self._names[k]] = v if keep_vars else v.detach()
        return out

    def _save_to_state_dict(self, destination, prefix, keep_vars):      # nested in a parent module: parent.state_dict() works
        self.state_dict(destination=destination, prefix=prefix, keep_vars=keep_vars)

    def load_state_dict(self, sd, strict=True, assign=False, prefix=""):
        missing = []
        known = set(self._names.values())
        unexpected = [k for k in sd if k.startswith(prefix) and k[len(prefix):] not in known and not k[len(prefix):].startswith("quantize.")]
        with torch.no_grad():
            for key, name in self._names.items():
                if prefix + name in sd:
                    self._parameters[key].copy_(sd[prefix + name])
                else:
                    missing.append(name)
        if strict and (missing or unexpected):
            raise KeyError(f"MAGVITv2.load_state_dict: missing keys {missing[:5]}{'...' if len(missing) > 5 else ''}, "
                           f"unexpected keys {unexpected[:5]}{'...' if len(unexpected) > 5 else ''}")
        self._versions = None            # in-place copies bump _version, but be explicit: the engine reloads on next use
        return missing, unexpected

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        m, u = self.load_state_dict(state_dict, strict=False, prefix=prefix)
        missing_keys.extend(prefix + k for k in m)
        unexpected_keys.extend(u)

    @classmethod
    def from_pretrained(cls, path, **kw):
        """Weights of `showlab/magvitv2` as laid out by save_pretrained (inference_t2i.py:61): single file or sharded."""
        from . import checkpoint
        model = cls()
        sd = checkpoint.read_state_dict(path)
        missing, _ = model.load_state_dict(sd, strict=False)
        if missing:
            raise KeyError(f"MAGVITv2.from_pretrained({path}): missing keys {missing[:5]}{'...' if len(missing) > 5 else ''}")
        return model

    def save_pretrained(self, path, max_shard_bytes: int = 5 << 30):
        from . import checkpoint
        if not self._parameters:
            raise _lib.ShowoError("MAGVITv2.save_pretrained: the weights were streamed into the engine (materialize=False); "
                                  "there are no torch parameters to write")
        checkpoint.write_checkpoint(path, self.state_dict(), {"_class_name": "MAGVITv2"}, max_shard_bytes)

    @property
    def device(self):
        for p in self._parameters.values():
            return p.device
        return torch.device("cuda", torch.cuda.current_device())

    def load_weights(self, weights: Dict[str, torch.Tensor], device=None):
        lib = _lib.require_gpu()
        device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if self._engine is None:
            h = C.c_void_p()
            _lib.check(lib.magvit_engine_create(device.index or 0, C.byref(h)), "magvit_engine_create")
            self._engine = h
        for name, t in weights.items():
            if name.startswith("quantize."):
                continue
            t = t.detach().float().contiguous()
            _lib.check(lib.magvit_load_weight(self._engine, name.encode(), _lib.ptr(t), t.numel(), int(t.is_cuda)),
                       f"magvit_load_weight({name})")
        _lib.check(lib.magvit_weights_complete(self._engine), "magvit_weights_complete")
        self._streamed = True
        return self

    def _sync(self):
        if not self._parameters:
            if not self._streamed:
                raise _lib.ShowoError("no weights: call load_weights()")
            return self._engine
        dev = self.device
        if dev.type != "cuda":
            raise _lib.ShowoError("MAGVITv2 must live on a CUDA (B200) device: show-o_b200 has no CPU fallback")
        versions = (str(dev),) + tuple((p._version, p.data_ptr()) for p in self._parameters.values())   # device moves re-upload too
        if self._engine is None or versions != self._versions:
            with torch.cuda.device(dev):
                self.load_weights(self.state_dict(), device=dev)
            self._versions = versions
        return self._engine

    def __del__(self):
        try:
            if self._engine is not None:
                _lib.load().magvit_engine_destroy(self._engine)
        except Exception:
            pass

    def kernel_launches(self) -> int:
        return int(_lib.load().magvit_kernel_launches(self._engine)) if self._engine is not None else 0

    @torch.no_grad()
    def get_code(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """modeling_magvitv2.py:423-427: [B,3,R,R] fp32 in [-1,1] -> LongTensor [B, (R/16)^2]."""
        lib = _lib.require_gpu()
        eng = self._sync()
        x = pixel_values.float().contiguous()
        B, _, R, R2 = x.shape
        assert R == R2
        out = torch.empty(B, (R // 16) ** 2, dtype=torch.int64, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.magvit_get_code(eng, _lib.ptr(x), B, R, _lib.ptr(out), _lib.current_stream_ptr()),
                       "magvit_get_code")
        return out

    def _grid(self, codebook_indices, shape):
        if shape is None:
            s = int(math.sqrt(codebook_indices.shape[-1]))
            return s, s
        return int(shape[0]), int(shape[1])

    @torch.no_grad()
    def decode_code(self, codebook_indices: torch.Tensor, shape=None) -> torch.Tensor:
        """modeling_magvitv2.py:429-433: LongTensor [B, h*w] -> fp32 [B,3,16h,16w]."""
        lib = _lib.require_gpu()
        eng = self._sync()
        ids = codebook_indices.to(torch.int64).contiguous()
        h, w = self._grid(ids, shape)
        B = ids.shape[0]
        out = torch.empty(B, 3, 16 * h, 16 * w, dtype=torch.float32, device=ids.device)
        with torch.cuda.device(ids.device):
            _lib.check(lib.magvit_decode_code(eng, _lib.ptr(ids), B, h, w, _lib.ptr(out), _lib.current_stream_ptr()),
                       "magvit_decode_code")
        return out

    @torch.no_grad()
    def decode_code_uint8(self, codebook_indices: torch.Tensor, shape=None) -> torch.Tensor:
        """decode_code fused with the caller's post-processing (inference_t2i.py:338-341) -> uint8 [B,16h,16w,3]."""
        lib = _lib.require_gpu()
        eng = self._sync()
        ids = codebook_indices.to(torch.int64).contiguous()
        h, w = self._grid(ids, shape)
        B = ids.shape[0]
        out = torch.empty(B, 16 * h, 16 * w, 3, dtype=torch.uint8, device=ids.device)
        with torch.cuda.device(ids.device):
            _lib.check(lib.magvit_decode_code_u8(eng, _lib.ptr(ids), B, h, w, _lib.ptr(out),
                                                 _lib.current_stream_ptr()), "magvit_decode_code_u8")
        return out

    # ------------------------------------------------------------------ fp32 verification path (parity tests only)
    @torch.no_grad()
    def get_code_fp32(self, pixel_values: torch.Tensor, return_z: bool = False):
        """get_code on the engine's fp32 verification path (fp32 NCHW activations and weights, CUDA cores only): the codes, and with
        `return_z` the quantizer's pre-sign values [B, 13, R/16, R/16] every code bit is the sign of."""
        lib = _lib.require_gpu()
        eng = self._sync()
        x = pixel_values.float().contiguous()
        B, _, R, R2 = x.shape
        assert R == R2
        out = torch.empty(B, (R // 16) ** 2, dtype=torch.int64, device=x.device)
        z = torch.empty(B, 13, R // 16, R // 16, dtype=torch.float32, device=x.device) if return_z else None
        with torch.cuda.device(x.device):
            _lib.check(lib.magvit_get_code_fp32(eng, _lib.ptr(x), B, R, _lib.ptr(out), _lib.ptr(z), _lib.current_stream_ptr()),
                       "magvit_get_code_fp32")
        return (out, z) if return_z else out

    @torch.no_grad()
    def decode_code_fp32(self, codebook_indices: torch.Tensor, shape=None) -> torch.Tensor:
        """decode_code on the fp32 verification path."""
        lib = _lib.require_gpu()
        eng = self._sync()
        ids = codebook_indices.to(torch.int64).contiguous()
        h, w = self._grid(ids, shape)
        B = ids.shape[0]
        out = torch.empty(B, 3, 16 * h, 16 * w, dtype=torch.float32, device=ids.device)
        with torch.cuda.device(ids.device):
            _lib.check(lib.magvit_decode_code_fp32(eng, _lib.ptr(ids), B, h, w, _lib.ptr(out), _lib.current_stream_ptr()),
                       "magvit_decode_code_fp32")
        return out
