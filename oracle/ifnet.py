"""fp32 CPU restatement of RIFE 4.26-heavy (test infrastructure, see oracle/__init__.py).

Functional torch code over a plain ``{name: tensor}`` state dict -- no nn.Module, no
autocast, no caches -- following
  models/rife_426_heavy/IFNet_HDv3.py:28-177  (Head, ResConv, IFBlock, IFNet.forward)
  models/rife_426_heavy/warplayer.py:8-22     (backward warp, border, align_corners=True)
  models/rife.py:25-109                       (inference_ts, calc_flow, inference_ts_drba)
with the splat / DRM arithmetic delegated to the C oracle (oracle/drba_oracle.c).
Pinned against the reference modules by tests/golden/rife_golden.npz.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import cport


def _lrelu(x):
    return F.leaky_relu(x, 0.2)


def encode(sd, img):
    """Head.forward: IFNet_HDv3.py:37-47."""
    x = _lrelu(F.conv2d(img, sd["encode.cnn0.weight"], sd["encode.cnn0.bias"], 2, 1))
    x = _lrelu(F.conv2d(x, sd["encode.cnn1.weight"], sd["encode.cnn1.bias"], 1, 1))
    x = _lrelu(F.conv2d(x, sd["encode.cnn2.weight"], sd["encode.cnn2.bias"], 1, 1))
    return F.conv_transpose2d(x, sd["encode.cnn3.weight"], sd["encode.cnn3.bias"], 2, 1)


def ifblock(sd, name, x, flow, scale):
    """IFBlock.forward: IFNet_HDv3.py:84-96."""
    x = F.interpolate(x, scale_factor=1.0 / scale, mode="bilinear", align_corners=False)
    if flow is not None:
        flow = F.interpolate(flow, scale_factor=1.0 / scale, mode="bilinear", align_corners=False) * 1.0 / scale
        x = torch.cat((x, flow), 1)
    h = _lrelu(F.conv2d(x, sd[f"{name}.conv0.0.0.weight"], sd[f"{name}.conv0.0.0.bias"], 2, 1))
    h = _lrelu(F.conv2d(h, sd[f"{name}.conv0.1.0.weight"], sd[f"{name}.conv0.1.0.bias"], 2, 1))
    for i in range(8):  # ResConv: IFNet_HDv3.py:58-59
        p = f"{name}.convblock.{i}"
        h = _lrelu(F.conv2d(h, sd[p + ".conv.weight"], sd[p + ".conv.bias"], 1, 1) * sd[p + ".beta"] + h)
    tmp = F.conv_transpose2d(h, sd[f"{name}.lastconv.0.weight"], sd[f"{name}.lastconv.0.bias"], 2, 1)
    tmp = F.pixel_shuffle(tmp, 2)
    tmp = F.interpolate(tmp, scale_factor=scale, mode="bilinear", align_corners=False)
    return tmp[:, :4] * scale, tmp[:, 4:5], tmp[:, 5:]


def backwarp(x, flow):
    """warplayer.py:8-22 in pixel coordinates (C oracle)."""
    return torch.from_numpy(cport.backwarp(x.numpy(), flow.numpy(), "border"))


def ifnet_forward(sd, x, timestep, scale_list, f0=None, f1=None):
    """IFNet.forward: IFNet_HDv3.py:126-177 (inference branch); returns (merged, flow_list)."""
    img0, img1 = x[:, :3], x[:, 3:6]
    if not torch.is_tensor(timestep):
        timestep = (x[:, :1].clone() * 0 + 1) * timestep
    f0 = encode(sd, img0) if f0 is None else f0
    f1 = encode(sd, img1) if f1 is None else f1
    flow = mask = feat = None
    w0, w1 = img0, img1
    flows = []
    for i in range(5):
        if flow is None:
            flow, mask, feat = ifblock(sd, "block0", torch.cat((img0, img1, f0, f1, timestep), 1), None, scale_list[0])
        else:
            wf0, wf1 = backwarp(f0, flow[:, :2]), backwarp(f1, flow[:, 2:4])
            fd, mask, feat = ifblock(sd, f"block{i}", torch.cat((w0, w1, wf0, wf1, timestep, mask, feat), 1),
                                     flow, scale_list[i])
            flow = flow + fd
        flows.append(flow)
        w0, w1 = backwarp(img0, flow[:, :2]), backwarp(img1, flow[:, 2:4])
    m = torch.sigmoid(mask)
    return w0 * m + w1 * (1 - m), flows


class RIFEOracle:
    """models/rife.py:15-109 restated in fp32 (the CuPy path's splat precision)."""

    def __init__(self, state, scale=1.0):
        self.sd = {k: v.float() for k, v in state.items()}
        self.scale = scale
        self.scale_list = [16 / scale, 8 / scale, 4 / scale, 2 / scale, 1 / scale]
        self.pad_size = 64

    def encode(self, img):
        return encode(self.sd, img)

    def inference_ts(self, I0, I1, ts):
        out = []
        for t in ts:
            if t == 0:
                out.append(I0)
            elif t == 1:
                out.append(I1)
            else:
                out.append(ifnet_forward(self.sd, torch.cat((I0, I1), 1), float(t), self.scale_list)[0])
        return out

    def calc_flow(self, a, b, f0=None, f1=None):
        """rife.py:41-75."""
        timestep = (a[:, :1].clone() * 0 + 1) * 0.5
        f0 = encode(self.sd, a[:, :3]) if f0 is None else f0
        f1 = encode(self.sd, b[:, :3]) if f1 is None else f1
        flow, _, _ = ifblock(self.sd, "block0", torch.cat((a[:, :3], b[:, :3], f0, f1, timestep), 1), None,
                             self.scale_list[0])
        flow01 = torch.from_numpy(cport.rife_invert_flow(flow[:, :2].contiguous().numpy()))
        flow10 = torch.from_numpy(cport.rife_invert_flow(flow[:, 2:].contiguous().numpy()))
        return flow01, flow10, f0, f1

    def inference_ts_drba(self, I0, I1, I2, ts, reuse=None, linear=False):
        """rife.py:79-109."""
        flow10, flow01, f1, f0 = self.calc_flow(I1, I0) if not reuse else reuse
        if reuse is None:
            flow12, flow21, f1, f2 = self.calc_flow(I1, I2)
        else:
            flow12, flow21, f1, f2 = self.calc_flow(I1, I2, f0=reuse[2])
        out = []
        for t in ts:
            if t == 0:
                out.append(I0)
            elif t == 1:
                out.append(I1)
            elif t == 2:
                out.append(I2)
            elif 0 < t < 1:
                drm = cport.calc_drm_rife(1 - t, flow10.numpy(), flow12.numpy(), linear)
                out.append(ifnet_forward(self.sd, torch.cat((I1, I0), 1), torch.from_numpy(drm["drm_t1_t01"]),
                                         self.scale_list, f0=f1, f1=f0)[0])
            elif 1 < t < 2:
                drm = cport.calc_drm_rife(t - 1, flow10.numpy(), flow12.numpy(), linear)
                out.append(ifnet_forward(self.sd, torch.cat((I1, I2), 1), torch.from_numpy(drm["drm_t1_t12"]),
                                         self.scale_list, f0=f1, f1=f2)[0])
        return out, (flow21, flow12, f2, f1)


def psnr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    mse = float(np.mean((a - b) ** 2))
    return 99.0 if mse == 0 else float(10 * np.log10(1.0 / mse))
