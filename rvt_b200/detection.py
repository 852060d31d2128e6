"""Inference mirror of the step right after the backbone: ``YOLOPAFPN``
(models/detection/yolox_extension/models/yolo_pafpn.py:18-139), ``YOLOXHead`` (models/detection/yolox/models/yolo_head.py:21-290,
inference branch) and ``postprocess`` (models/detection/yolox/utils/boxes.py:32-76), on the CUDA library.

Same constructor arguments, ``forward`` signatures and ``state_dict`` keys (``lateral_conv0.conv.weight``, ``...bn.running_mean``,
``C3_p4.m.0.conv2.conv.weight``, ``cls_convs.0.1.conv.weight``, ``cls_preds.0.bias`` ...), so a released checkpoint's ``fpn.*`` /
``yolox_head.*`` entries load strictly.  Every BaseConv (Conv2d + BatchNorm2d(eval) + SiLU, network_blocks.py:29-51) is ONE
implicit-GEMM launch with the BatchNorm folded into the packed weight and a bias; ``th.cat`` along channels and the
nearest-exact upsample write channel slices of shared channels-last fp16 buffers.  Training (losses / SimOTA) is the reference's
own code path and is not rebuilt: ``forward`` with labels raises."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, packing
from .ops import device_guarded


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _ru(n, m):
    return (n + m - 1) // m * m


class _BN(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer('running_mean', torch.zeros(c))
        self.register_buffer('running_var', torch.ones(c))
        self.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long))
        self.eps = 1e-5


class _ConvW(nn.Module):
    def __init__(self, cin, cout, k, bias=False):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        if bias:
            self.bias = nn.Parameter(torch.zeros(cout))


class BaseConv(nn.Module):
    """parameters of network_blocks.py:29-51 (conv.weight, bn.*); act must be silu"""

    def __init__(self, in_channels, out_channels, ksize, stride, act='silu'):
        super().__init__()
        if act != 'silu':
            raise NotImplementedError(f'act={act}: only silu (every released config) is built')
        self.cin, self.cout, self.ksize, self.stride = in_channels, out_channels, ksize, stride
        self.conv = _ConvW(in_channels, out_channels, ksize)
        self.bn = _BN(out_channels)

    def folded(self, device):
        """(packed fp16 weight tiles, fp32 bias) with the eval-mode BatchNorm folded in"""
        s = (self.bn.weight / torch.sqrt(self.bn.running_var + self.bn.eps)).detach().float().to(device)
        w = self.conv.weight.detach().float().to(device) * s.view(-1, 1, 1, 1)
        b = (self.bn.bias.detach().float().to(device) - self.bn.running_mean.float().to(device) * s).contiguous()
        cp = _ru(self.cout, 16)
        if cp != self.cout:
            w = torch.cat([w, torch.zeros(cp - self.cout, *w.shape[1:], device=device)])
            b = torch.cat([b, torch.zeros(cp - self.cout, device=device)])
        bn = _lib.lib().rvt_tile_n(cp, self.cin * self.ksize ** 2)
        return packing.pack_conv_weight(w, channels_last_input=True, bn=bn), b.contiguous()


class Bottleneck(nn.Module):
    def __init__(self, cin, cout, shortcut, expansion, depthwise, act):
        super().__init__()
        hidden = int(cout * expansion)
        self.conv1 = BaseConv(cin, hidden, 1, 1, act)
        self.conv2 = BaseConv(hidden, cout, 3, 1, act)
        self.use_add = shortcut and cin == cout


class CSPLayer(nn.Module):
    def __init__(self, cin, cout, n=1, shortcut=True, expansion=0.5, depthwise=False, act='silu'):
        super().__init__()
        hidden = int(cout * expansion)
        self.hidden = hidden
        self.conv1 = BaseConv(cin, hidden, 1, 1, act)
        self.conv2 = BaseConv(cin, hidden, 1, 1, act)
        self.conv3 = BaseConv(2 * hidden, cout, 1, 1, act)
        self.m = nn.Sequential(*[Bottleneck(hidden, hidden, shortcut, 1.0, depthwise, act) for _ in range(n)])


class _Slice:
    """channel slice [c0, c0 + c) of a channels-last fp16 buffer [rows_padded, pitch] holding B x H x W pixels"""

    def __init__(self, buf, b, h, w, pitch, c0, c):
        self.buf, self.b, self.h, self.w, self.pitch, self.c0, self.c = buf, b, h, w, pitch, c0, c

    def ptr(self):
        return self.buf.data_ptr() + 2 * self.c0


def _new_buf(b, h, w, c, dev):
    rows = _ru(b * h * w, 128)
    buf = torch.empty(rows * c, dtype=torch.float16, device=dev)
    return _Slice(buf, b, h, w, c, 0, c)


class _Engine:
    """packed weights + launch helpers shared by the FPN and the head"""

    def __init__(self):
        self.cache: Dict[int, tuple] = {}

    def packed(self, conv: BaseConv, dev):
        key = (id(conv), conv.conv.weight._version, conv.bn.weight._version, conv.bn.running_var._version)
        hit = self.cache.get(id(conv))
        if hit is None or hit[0] != key:
            hit = (key, conv.folded(dev))
            self.cache[id(conv)] = hit
        return hit[1]

    def conv(self, conv: BaseConv, src: _Slice, dst: Optional[_Slice] = None) -> _Slice:
        dev = src.buf.device
        wp, bias = self.packed(conv, dev)
        k, s = conv.ksize, conv.stride
        pad = (k - 1) // 2
        ho, wo = (src.h + 2 * pad - k) // s + 1, (src.w + 2 * pad - k) // s + 1
        cp = _ru(conv.cout, 16)
        if dst is None:
            dst = _new_buf(src.b, ho, wo, cp, dev)
            dst.c = conv.cout
        assert src.c == conv.cin and dst.h == ho and dst.w == wo and dst.c0 % 8 == 0
        assert dst.c0 + cp <= dst.pitch, 'destination slice must have room for the padded channel count'
        _lib.check(_lib.lib().rvt_conv2d_nhwc_f16(src.ptr(), src.pitch, src.b, conv.cin, src.h, src.w, k, s, pad, ho, wo, cp,
                                                  _lib.ptr(wp), _lib.ptr(bias), 4, dst.ptr(), dst.pitch, _stream(src.buf)),
                   'conv2d_nhwc_f16')
        return dst

    @staticmethod
    def cast_in(x: torch.Tensor, dst: _Slice):
        """backbone feature (logical NCHW fp32, any strides) -> fp16 channel slice"""
        b, c, h, w = x.shape
        x = x.detach()
        if x.dtype != torch.float32:
            x = x.float()
        assert (b, h, w, c) == (dst.b, dst.h, dst.w, dst.c)
        sb, sc, sy, sx = x.stride()
        _lib.check(_lib.lib().rvt_cast_slice_f16(x.data_ptr(), sb, sy, sx, sc, b, h, w, c, dst.ptr(), dst.pitch, _stream(x)),
                   'cast_slice_f16')

    @staticmethod
    def upsample_into(src: _Slice, dst: _Slice):
        assert dst.h == 2 * src.h and dst.w == 2 * src.w and dst.c == src.c
        _lib.check(_lib.lib().rvt_upsample2_slice_f16(src.ptr(), src.pitch, src.b, src.h, src.w, src.c, dst.ptr(), dst.pitch,
                                                      _stream(src.buf)), 'upsample2_slice_f16')

    def csp(self, layer: CSPLayer, src: _Slice) -> _Slice:
        """CSPLayer.forward (network_blocks.py:135-141): cat(m(conv1(x)), conv2(x)) -> conv3; the cat is one buffer"""
        dev = src.buf.device
        hid = layer.hidden
        cat = _new_buf(src.b, src.h, src.w, 2 * _ru(hid, 16), dev)
        assert hid % 16 == 0, 'CSP hidden width must be a multiple of 16'
        x1 = self.conv(layer.conv1, src)
        for bt in layer.m:
            y = self.conv(bt.conv1, x1)
            last = bt is layer.m[-1]
            y = self.conv(bt.conv2, y, _Slice(cat.buf, src.b, src.h, src.w, cat.pitch, 0, hid) if (last and not bt.use_add) else None)
            if bt.use_add:
                raise NotImplementedError('Bottleneck shortcut (unused by YOLOPAFPN: shortcut=False)')
            x1 = y
        if len(layer.m) == 0:
            raise NotImplementedError('CSPLayer with n = 0')
        self.conv(layer.conv2, src, _Slice(cat.buf, src.b, src.h, src.w, cat.pitch, hid, hid))
        return self.conv(layer.conv3, _Slice(cat.buf, src.b, src.h, src.w, cat.pitch, 0, 2 * hid))


class YOLOPAFPN(nn.Module):
    """yolo_pafpn.py:18-139"""

    def __init__(self, depth: float = 1.0, in_stages: Tuple[int, ...] = (2, 3, 4), in_channels: Tuple[int, ...] = (256, 512, 1024),
                 depthwise: bool = False, act: str = 'silu', compile_cfg: Optional[Dict] = None):
        super().__init__()
        assert len(in_stages) == len(in_channels) == 3
        if depthwise:
            raise NotImplementedError('depthwise=True (DWConv) is not built; released configs use False')
        self.in_features, self.in_channels = tuple(in_stages), tuple(in_channels)
        c0, c1, c2 = in_channels
        n = round(3 * depth)
        self.lateral_conv0 = BaseConv(c2, c1, 1, 1, act)
        self.C3_p4 = CSPLayer(2 * c1, c1, n, False, act=act)
        self.reduce_conv1 = BaseConv(c1, c0, 1, 1, act)
        self.C3_p3 = CSPLayer(2 * c0, c0, n, False, act=act)
        self.bu_conv2 = BaseConv(c0, c0, 3, 2, act)
        self.C3_n3 = CSPLayer(2 * c0, c1, n, False, act=act)
        self.bu_conv1 = BaseConv(c1, c1, 3, 2, act)
        self.C3_n4 = CSPLayer(2 * c1, c2, n, False, act=act)
        self._eng = _Engine()

    @torch.no_grad()
    def forward_slices(self, feats: Dict[int, torch.Tensor]):
        """-> (pan_out2, pan_out1, pan_out0) as channels-last fp16 slices (what YOLOXHead consumes)"""
        x2, x1, x0 = (feats[f] for f in self.in_features)
        if not x0.is_cuda:
            raise RuntimeError('rvt_b200.detection runs on CUDA (sm_100a) only; there is no CPU fallback')
        dev = x0.device
        c0, c1, c2 = self.in_channels
        e = self._eng
        b = x0.shape[0]
        (h2, w2), (h1, w1), (h0, w0) = x2.shape[-2:], x1.shape[-2:], x0.shape[-2:]
        with torch.cuda.device(dev):
            s0 = _new_buf(b, h0, w0, c2, dev)
            e.cast_in(x0, s0)
            cat_n4 = _new_buf(b, h0, w0, 2 * c1, dev)                       # [bu_conv1(pan_out1) | fpn_out0]
            fpn_out0 = e.conv(self.lateral_conv0, s0, _Slice(cat_n4.buf, b, h0, w0, cat_n4.pitch, c1, c1))
            cat_p4 = _new_buf(b, h1, w1, 2 * c1, dev)                       # [up(fpn_out0) | x1]
            e.upsample_into(fpn_out0, _Slice(cat_p4.buf, b, h1, w1, cat_p4.pitch, 0, c1))
            e.cast_in(x1, _Slice(cat_p4.buf, b, h1, w1, cat_p4.pitch, c1, c1))
            f_out0 = e.csp(self.C3_p4, cat_p4)
            cat_n3 = _new_buf(b, h1, w1, 2 * c0, dev)                       # [bu_conv2(pan_out2) | fpn_out1]
            fpn_out1 = e.conv(self.reduce_conv1, f_out0, _Slice(cat_n3.buf, b, h1, w1, cat_n3.pitch, c0, c0))
            cat_p3 = _new_buf(b, h2, w2, 2 * c0, dev)                       # [up(fpn_out1) | x2]
            e.upsample_into(fpn_out1, _Slice(cat_p3.buf, b, h2, w2, cat_p3.pitch, 0, c0))
            e.cast_in(x2, _Slice(cat_p3.buf, b, h2, w2, cat_p3.pitch, c0, c0))
            pan_out2 = e.csp(self.C3_p3, cat_p3)
            e.conv(self.bu_conv2, pan_out2, _Slice(cat_n3.buf, b, h1, w1, cat_n3.pitch, 0, c0))
            pan_out1 = e.csp(self.C3_n3, cat_n3)
            e.conv(self.bu_conv1, pan_out1, _Slice(cat_n4.buf, b, h0, w0, cat_n4.pitch, 0, c1))
            pan_out0 = e.csp(self.C3_n4, cat_n4)
        return pan_out2, pan_out1, pan_out0

    def forward(self, input: Dict[int, torch.Tensor]):
        """reference API: tuple of NCHW fp32 feature maps (materialised from the fp16 slices)"""
        outs = []
        for sl in self.forward_slices(input):
            t = sl.buf.view(-1, sl.pitch)[:sl.b * sl.h * sl.w, sl.c0:sl.c0 + sl.c].float()
            outs.append(t.reshape(sl.b, sl.h, sl.w, sl.c).permute(0, 3, 1, 2))
        return tuple(outs)


class _Pred(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, 1, 1))
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        self.bias = nn.Parameter(torch.zeros(cout))


class YOLOXHead(nn.Module):
    """yolo_head.py:21-290, inference branch (decode_in_inference=True)"""

    def __init__(self, num_classes=80, strides=(8, 16, 32), in_channels=(256, 512, 1024), act='silu', depthwise=False,
                 compile_cfg: Optional[Dict] = None):
        super().__init__()
        if depthwise:
            raise NotImplementedError('depthwise=True (DWConv) is not built; released configs use False')
        import math
        self.num_classes, self.strides = num_classes, tuple(strides)
        self.decode_in_inference = True
        hidden = int(256 * in_channels[-1] / 1024)
        self.hidden = hidden
        self.stems = nn.ModuleList([BaseConv(c, hidden, 1, 1, act) for c in in_channels])
        self.cls_convs = nn.ModuleList([nn.Sequential(BaseConv(hidden, hidden, 3, 1, act), BaseConv(hidden, hidden, 3, 1, act))
                                        for _ in in_channels])
        self.reg_convs = nn.ModuleList([nn.Sequential(BaseConv(hidden, hidden, 3, 1, act), BaseConv(hidden, hidden, 3, 1, act))
                                        for _ in in_channels])
        self.cls_preds = nn.ModuleList([_Pred(hidden, num_classes) for _ in in_channels])
        self.reg_preds = nn.ModuleList([_Pred(hidden, 4) for _ in in_channels])
        self.obj_preds = nn.ModuleList([_Pred(hidden, 1) for _ in in_channels])
        prior = -math.log((1 - 0.01) / 0.01)                                 # initialize_biases(prior_prob=0.01), :152-163
        with torch.no_grad():
            for m in list(self.cls_preds) + list(self.obj_preds):
                m.bias.fill_(prior)
        self._eng = _Engine()
        self._pred_cache = {}

    def _pred_packed(self, k, dev):
        """1x1 prediction convs as two GEMMs per level: [reg(4) | obj(1)] from reg_feat, [cls] from cls_feat (N padded to 16)"""
        ps = (self.reg_preds[k], self.obj_preds[k], self.cls_preds[k])
        key = tuple(p.weight._version for p in ps) + tuple(p.bias._version for p in ps)
        hit = self._pred_cache.get(k)
        if hit is None or hit[0] != key or hit[1][0].device != dev:
            def pack(ws, bs):
                w = torch.cat([x.detach().float().to(dev).reshape(x.shape[0], -1) for x in ws])
                b = torch.cat([x.detach().float().to(dev) for x in bs])
                n = _ru(w.shape[0], 16)
                wpad = torch.zeros(n, w.shape[1], device=dev)
                wpad[:w.shape[0]] = w
                bpad = torch.zeros(n, device=dev)
                bpad[:b.shape[0]] = b
                return packing.pack_linear_weight(wpad, _lib.lib().rvt_tile_n(n, w.shape[1])), bpad.contiguous(), n
            hit = (key, pack([ps[0].weight, ps[1].weight], [ps[0].bias, ps[1].bias]) + pack([ps[2].weight], [ps[2].bias]))
            self._pred_cache[k] = hit
        return hit[1]

    @torch.no_grad()
    def forward(self, xin, labels=None):
        if labels is not None:
            raise NotImplementedError('YOLOXHead training (SimOTA losses) is the reference\'s own path; only inference is built')
        L = _lib.lib()
        e = self._eng
        if isinstance(xin[0], torch.Tensor):                               # reference API: NCHW tensors -> slices
            sl = []
            for x in xin:
                s = _new_buf(x.shape[0], x.shape[2], x.shape[3], x.shape[1], x.device)
                with torch.cuda.device(x.device):
                    e.cast_in(x, s)
                sl.append(s)
            xin = sl
        dev = xin[0].buf.device
        b = xin[0].b
        a_total = sum(s.h * s.w for s in xin)
        nc = self.num_classes
        out = torch.empty((b, a_total, 5 + nc), dtype=torch.float32, device=dev)
        a0 = 0
        with torch.cuda.device(dev):
            for k, (x, stride) in enumerate(zip(xin, self.strides)):
                f = e.conv(self.stems[k], x)
                cls_feat = e.conv(self.cls_convs[k][1], e.conv(self.cls_convs[k][0], f))
                reg_feat = e.conv(self.reg_convs[k][1], e.conv(self.reg_convs[k][0], f))
                w_ro, b_ro, n_ro, w_c, b_c, n_c = self._pred_packed(k, dev)
                rows = _ru(b * x.h * x.w, 128)
                regobj = torch.empty(rows * n_ro, dtype=torch.float16, device=dev)
                cls = torch.empty(rows * n_c, dtype=torch.float16, device=dev)
                st = _stream(out)
                _lib.check(L.rvt_conv2d_nhwc_f16(reg_feat.ptr(), reg_feat.pitch, b, self.hidden, x.h, x.w, 1, 1, 0, x.h, x.w, n_ro,
                                                 _lib.ptr(w_ro), _lib.ptr(b_ro), 0, _lib.ptr(regobj), n_ro, st), 'pred reg/obj')
                _lib.check(L.rvt_conv2d_nhwc_f16(cls_feat.ptr(), cls_feat.pitch, b, self.hidden, x.h, x.w, 1, 1, 0, x.h, x.w, n_c,
                                                 _lib.ptr(w_c), _lib.ptr(b_c), 0, _lib.ptr(cls), n_c, st), 'pred cls')
                _lib.check(L.rvt_yolox_decode(_lib.ptr(regobj), n_ro, _lib.ptr(cls), n_c, b, x.h, x.w, nc, float(stride), a0, a_total,
                                              _lib.ptr(out), st), 'yolox_decode')
                a0 += x.h * x.w
        self.hw = [(s.h, s.w) for s in xin]
        return out, None


@device_guarded
def postprocess(prediction: torch.Tensor, num_classes: int, conf_thre: float = 0.7, nms_thre: float = 0.45,
                class_agnostic: bool = False) -> List[Optional[torch.Tensor]]:
    """boxes.py:32-76: list (per image) of [n, 7] detections (x1, y1, x2, y2, obj_conf, class_conf, class_pred) after
    confidence filtering and per-class NMS, in score order; None for an image without detections."""
    if class_agnostic:
        raise NotImplementedError('class_agnostic=True is not built (the harness calls postprocess without it)')
    if not prediction.is_cuda:
        raise RuntimeError('rvt_b200.detection.postprocess runs on CUDA (sm_100a) only; there is no CPU fallback')
    pred = prediction.detach().float().contiguous()
    b, a, ch = pred.shape
    assert ch == 5 + num_classes
    det = torch.empty((b, a, 7), dtype=torch.float32, device=pred.device)
    counts = torch.zeros(b, dtype=torch.int32, device=pred.device)
    _lib.check(_lib.lib().rvt_yolox_postprocess(_lib.ptr(pred), b, a, num_classes, float(conf_thre), float(nms_thre), _lib.ptr(det),
                                                _lib.ptr(counts), _stream(pred)), 'yolox_postprocess')
    n = counts.cpu().tolist()                       # dynamic output shapes, as in the reference: one device->host read
    return [det[i, :n[i]] if n[i] > 0 else None for i in range(b)]


# ---------------------------------------------------------------------------------------------------------------------
# builders / detector glue (yolox_extension/models/build.py:9-28, detector.py:18-72)
# ---------------------------------------------------------------------------------------------------------------------
def _plain(cfg):
    if hasattr(cfg, 'items'):
        return {k: _plain(v) for k, v in cfg.items()}
    if isinstance(cfg, (list, tuple)):
        return [_plain(v) for v in cfg]
    return cfg


def build_yolox_head(head_cfg, in_channels, strides):
    d = _plain(head_cfg)
    d.pop('name', None)
    d.pop('version', None)
    d['in_channels'], d['strides'] = tuple(in_channels), tuple(strides)
    d['compile_cfg'] = d.pop('compile', None)
    return YOLOXHead(**d)


def build_yolox_fpn(fpn_cfg, in_channels):
    d = _plain(fpn_cfg)
    name = d.pop('name')
    if name not in ('PAFPN', 'pafpn'):
        raise NotImplementedError(name)
    d['in_channels'] = tuple(in_channels)
    d['in_stages'] = tuple(d['in_stages'])
    d['compile_cfg'] = d.pop('compile', None)
    return YOLOPAFPN(**d)


class YoloXDetector(nn.Module):
    """detector.py:18-72 for inference: ``backbone.*`` / ``fpn.*`` / ``yolox_head.*`` state_dict keys as in the reference."""

    def __init__(self, model_cfg):
        super().__init__()
        from .backbone import _cfg, build_recurrent_backbone
        fpn_cfg, head_cfg = _cfg(model_cfg, 'fpn'), _cfg(model_cfg, 'head')
        self.backbone = build_recurrent_backbone(_cfg(model_cfg, 'backbone'))
        in_stages = tuple(_cfg(fpn_cfg, 'in_stages'))
        in_channels = self.backbone.get_stage_dims(in_stages)
        self.fpn = build_yolox_fpn(fpn_cfg, in_channels=in_channels)
        self.yolox_head = build_yolox_head(head_cfg, in_channels=in_channels, strides=self.backbone.get_strides(in_stages))

    def forward_backbone(self, x, previous_states=None, token_mask=None):
        return self.backbone(x, previous_states, token_mask)

    def forward_detect(self, backbone_features, targets=None):
        if targets is not None or self.training:
            raise NotImplementedError('training of the detection head is the reference\'s own path; call .eval()')
        return self.yolox_head(self.fpn.forward_slices(backbone_features))

    def forward(self, x, previous_states=None, retrieve_detections: bool = True, targets=None):
        feats, states = self.forward_backbone(x, previous_states)
        if not retrieve_detections:
            assert targets is None
            return None, None, states
        outputs, losses = self.forward_detect(feats, targets)
        return outputs, losses, states
