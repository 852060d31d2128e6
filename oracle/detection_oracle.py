"""ORACLE — TEST INFRASTRUCTURE ONLY; never imported by rvt_b200/.

CPU fp32 restatement (plain functions over the reference's state_dict keys) of the detection step that follows the backbone
(SURVEY.md §8 f2), pinned against the imported reference by oracle/make_golden.py (tests/golden/det_*.npz):

  pafpn_forward   models/detection/yolox_extension/models/yolo_pafpn.py:109-139  (+ CSPLayer / Bottleneck / BaseConv,
                  models/detection/yolox/models/network_blocks.py:29-141; BatchNorm in eval mode)
  head_forward    models/detection/yolox/models/yolo_head.py:165-290 (inference branch + decode_outputs)
  postprocess     models/detection/yolox/utils/boxes.py:32-76 (confidence filter + torchvision batched_nms semantics)
"""
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def base_conv(x, p, pre, stride=1):
    """BaseConv.forward (network_blocks.py:50-51): act(bn(conv(x))), 'same' padding, BatchNorm2d eval, SiLU"""
    w = p[pre + 'conv.weight']
    y = F.conv2d(x, w, None, stride=stride, padding=(w.shape[-1] - 1) // 2)
    y = F.batch_norm(y, p[pre + 'bn.running_mean'], p[pre + 'bn.running_var'], p[pre + 'bn.weight'], p[pre + 'bn.bias'], False, 0.0, 1e-5)
    return y * torch.sigmoid(y)


def csp_layer(x, p, pre, n):
    """CSPLayer.forward (network_blocks.py:135-141) with shortcut=False bottlenecks (yolo_pafpn.py:55-62)"""
    x1 = base_conv(x, p, pre + 'conv1.')
    x2 = base_conv(x, p, pre + 'conv2.')
    for i in range(n):
        x1 = base_conv(base_conv(x1, p, f'{pre}m.{i}.conv1.'), p, f'{pre}m.{i}.conv2.')
    return base_conv(torch.cat((x1, x2), dim=1), p, pre + 'conv3.')


def pafpn_forward(feats: Dict[int, torch.Tensor], p, in_stages=(2, 3, 4), depth=0.67):
    n = round(3 * depth)
    x2, x1, x0 = (feats[s] for s in in_stages)
    up = lambda t: F.interpolate(t, scale_factor=2, mode='nearest-exact')
    fpn_out0 = base_conv(x0, p, 'lateral_conv0.')
    f_out0 = csp_layer(torch.cat([up(fpn_out0), x1], 1), p, 'C3_p4.', n)
    fpn_out1 = base_conv(f_out0, p, 'reduce_conv1.')
    pan_out2 = csp_layer(torch.cat([up(fpn_out1), x2], 1), p, 'C3_p3.', n)
    p_out1 = torch.cat([base_conv(pan_out2, p, 'bu_conv2.', 2), fpn_out1], 1)
    pan_out1 = csp_layer(p_out1, p, 'C3_n3.', n)
    p_out0 = torch.cat([base_conv(pan_out1, p, 'bu_conv1.', 2), fpn_out0], 1)
    pan_out0 = csp_layer(p_out0, p, 'C3_n4.', n)
    return pan_out2, pan_out1, pan_out0


def head_forward(xin, p, strides=(8, 16, 32)):
    """YOLOXHead.forward in eval mode -> decoded [B, n_anchors, 5 + nc] (yolo_head.py:176-232, 271-290)"""
    outs = []
    grids, strs = [], []
    for k, (x, s) in enumerate(zip(xin, strides)):
        f = base_conv(x, p, f'stems.{k}.')
        cf = base_conv(base_conv(f, p, f'cls_convs.{k}.0.'), p, f'cls_convs.{k}.1.')
        rf = base_conv(base_conv(f, p, f'reg_convs.{k}.0.'), p, f'reg_convs.{k}.1.')
        cls = F.conv2d(cf, p[f'cls_preds.{k}.weight'], p[f'cls_preds.{k}.bias'])
        reg = F.conv2d(rf, p[f'reg_preds.{k}.weight'], p[f'reg_preds.{k}.bias'])
        obj = F.conv2d(rf, p[f'obj_preds.{k}.weight'], p[f'obj_preds.{k}.bias'])
        o = torch.cat([reg, obj.sigmoid(), cls.sigmoid()], 1)
        h, w = o.shape[-2:]
        outs.append(o.flatten(start_dim=2))
        yv, xv = torch.meshgrid([torch.arange(h, dtype=o.dtype), torch.arange(w, dtype=o.dtype)], indexing='ij')
        grids.append(torch.stack((xv, yv), 2).view(1, -1, 2))
        strs.append(torch.full((1, h * w, 1), float(s), dtype=o.dtype))
    out = torch.cat(outs, dim=2).permute(0, 2, 1)
    g, st = torch.cat(grids, 1), torch.cat(strs, 1)
    return torch.cat([(out[..., 0:2] + g) * st, torch.exp(out[..., 2:4]) * st, out[..., 4:]], dim=-1)


def _iou(a, b):
    iw = max(min(a[2], b[2]) - max(a[0], b[0]), 0.0)
    ih = max(min(a[3], b[3]) - max(a[1], b[1]), 0.0)
    inter = iw * ih
    return inter / ((a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter)


def postprocess(prediction: np.ndarray, num_classes: int, conf_thre=0.7, nms_thre=0.45) -> List[Optional[np.ndarray]]:
    """numpy float32 restatement: corners, class max / argmax, score filter, score-sorted greedy per-class NMS"""
    out = []
    pred = np.asarray(prediction, dtype=np.float32)
    for img in pred:
        box = np.stack([img[:, 0] - img[:, 2] / np.float32(2), img[:, 1] - img[:, 3] / np.float32(2),
                        img[:, 0] + img[:, 2] / np.float32(2), img[:, 1] + img[:, 3] / np.float32(2)], 1).astype(np.float32)
        cls = img[:, 5:5 + num_classes]
        cconf, cpred = cls.max(1), cls.argmax(1)
        score = (img[:, 4] * cconf).astype(np.float32)
        keep = np.nonzero(score >= np.float32(conf_thre))[0]
        if keep.size == 0:
            out.append(None)
            continue
        order = keep[np.argsort(-score[keep], kind='stable')]
        sel = []
        for i in order:
            if all(cpred[i] != cpred[j] or np.float32(_iou(box[i], box[j])) <= np.float32(nms_thre) for j in sel):
                sel.append(i)
        sel = np.array(sel)
        out.append(np.concatenate([box[sel], img[sel, 4:5], cconf[sel, None], cpred[sel, None].astype(np.float32)], 1))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# deterministic synthetic parameters (BatchNorm statistics randomised so the fold is actually exercised)
# ---------------------------------------------------------------------------------------------------------------------
def synth_state(shapes: Dict[str, Tuple[int, ...]], seed: int) -> Dict[str, torch.Tensor]:
    import zlib
    out = {}
    for k, shape in shapes.items():
        rs = np.random.RandomState((seed * 1000003 + zlib.crc32(k.encode())) % (2 ** 31))     # independent of the key order
        n = int(np.prod(shape)) if len(shape) else 1
        if k.endswith('num_batches_tracked'):
            out[k] = torch.tensor(7, dtype=torch.long)
            continue
        if k.endswith('running_var'):
            v = rs.uniform(0.5, 1.5, n)
        elif k.endswith('running_mean'):
            v = rs.normal(0, 0.2, n)
        elif k.endswith('bn.weight'):
            v = rs.uniform(0.7, 1.3, n)
        elif k.endswith('bias'):
            v = rs.normal(0, 0.1, n)
        else:
            fan_in = int(np.prod(shape[1:]))
            v = rs.normal(0, 1.4 / np.sqrt(fan_in), n)
        out[k] = torch.from_numpy(v.astype(np.float32).reshape(shape))
    return out


def synth_predictions(seed: int, batch: int, anchors: int, num_classes: int, hw=(384, 640)) -> np.ndarray:
    """decoded-head-like tensor with clusters of overlapping boxes and a realistic share of confident anchors"""
    rs = np.random.RandomState(seed)
    pred = np.zeros((batch, anchors, 5 + num_classes), np.float32)
    for b in range(batch):
        n_obj = rs.randint(3, 9)
        centres = rs.uniform([40, 40], [hw[1] - 40, hw[0] - 40], (n_obj, 2))
        sizes = rs.uniform(20, 120, (n_obj, 2))
        which = rs.randint(0, n_obj, anchors)
        pred[b, :, 0:2] = centres[which] + rs.normal(0, 6, (anchors, 2))
        pred[b, :, 2:4] = np.maximum(sizes[which] * rs.uniform(0.8, 1.25, (anchors, 2)), 2)
        conf = rs.uniform(size=anchors) < 0.06
        pred[b, :, 4] = np.where(conf, rs.uniform(0.3, 1.0, anchors), rs.uniform(0, 0.05, anchors))
        pred[b, :, 5:] = rs.uniform(0, 1, (anchors, num_classes))
    pred[batch - 1, :, 4] = 0.001                                  # one image without any detection
    return pred
