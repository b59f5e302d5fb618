"""Weight handling for the RIFE 4.26-heavy IFNet (models/rife.py:19-20, tools.py:83-88).

* ``load_ifnet_state(dir)`` reads the reference's ``flownet.pkl`` (a state dict whose keys
  carry a ``module.`` prefix; 40 unused ``teacher.*``/``caltime.*`` entries are dropped).
* ``synth_ifnet_state(seed)`` builds a state dict of the same names and shapes from a
  seeded CPU generator.  It is used wherever the reference checkpoint is not available
  (the GPU box receives only this repository) so that kernels, parity tests and
  throughput runs see identical weights on every machine.
"""
import os

import torch

# name -> shape, in the order of the reference module tree
# (models/rife_426_heavy/IFNet_HDv3.py:28-120)
_BLOCKS = [("block0", 7 + 32, 192), ("block1", 52, 128), ("block2", 52, 96), ("block3", 52, 64), ("block4", 52, 32)]


def ifnet_param_shapes():
    shapes = []
    for name, cin, c in _BLOCKS:
        shapes += [(f"{name}.conv0.0.0.weight", (c // 2, cin, 3, 3)), (f"{name}.conv0.0.0.bias", (c // 2,)),
                   (f"{name}.conv0.1.0.weight", (c, c // 2, 3, 3)), (f"{name}.conv0.1.0.bias", (c,))]
        for i in range(8):
            shapes += [(f"{name}.convblock.{i}.beta", (1, c, 1, 1)),
                       (f"{name}.convblock.{i}.conv.weight", (c, c, 3, 3)),
                       (f"{name}.convblock.{i}.conv.bias", (c,))]
        shapes += [(f"{name}.lastconv.0.weight", (c, 52, 4, 4)), (f"{name}.lastconv.0.bias", (52,))]
    shapes += [("encode.cnn0.weight", (16, 3, 3, 3)), ("encode.cnn0.bias", (16,)),
               ("encode.cnn1.weight", (16, 16, 3, 3)), ("encode.cnn1.bias", (16,)),
               ("encode.cnn2.weight", (16, 16, 3, 3)), ("encode.cnn2.bias", (16,)),
               ("encode.cnn3.weight", (16, 16, 4, 4)), ("encode.cnn3.bias", (16,))]
    return shapes


def synth_ifnet_state(seed=0):
    """Deterministic stand-in weights: He-style fan-in scaling so activations stay O(1),
    residual gains (beta) around 0.3 like the trained checkpoint, and a small last-layer
    gain so synthetic flows stay within a few pixels per pyramid level."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000 + int(seed))
    state = {}
    for name, shape in ifnet_param_shapes():
        if name.endswith("beta"):
            v = 0.2 + 0.2 * torch.rand(shape, generator=g)
        elif name.endswith("bias"):
            v = 0.05 * torch.randn(shape, generator=g)
        else:
            if "lastconv" in name or name == "encode.cnn3.weight":  # ConvTranspose2d: [Cin, Cout, 4, 4]
                fan_in = shape[0] * 4
                gain = 0.5 if "lastconv" in name else 1.0
            else:
                fan_in = shape[1] * shape[2] * shape[3]
                gain = 1.0
            v = torch.randn(shape, generator=g) * (gain * (1.6 / fan_in) ** 0.5)
        state[name] = v.float()
    return state


def load_ifnet_state(weights_dir):
    """Read ``flownet.pkl`` the way models/rife.py:19 + tools.py:83 (`convert`) do."""
    raw = torch.load(os.path.join(weights_dir, "flownet.pkl"), map_location="cpu")
    want = dict(ifnet_param_shapes())
    state = {}
    for k, v in raw.items():
        if "module." not in k:
            continue
        k2 = k.replace("module.", "")
        if k2 in want:
            state[k2] = v.detach().float().contiguous()
    missing = [k for k in want if k not in state]
    if missing:
        raise KeyError(f"flownet.pkl is missing {len(missing)} tensors, e.g. {missing[:3]}")
    return state


def find_rife_weights(explicit=None):
    """Locate a RIFE checkpoint directory; returns None when only synthetic weights are possible."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cands = [explicit, os.environ.get("DRBA_RIFE_WEIGHTS"),
             "weights/train_log_rife_426_heavy",
             os.path.join(here, "weights/train_log_rife_426_heavy"),
             os.path.join(here, "oracle/_ref/weights/train_log_rife_426_heavy"),
             os.path.join(here, "baseline/_ref/weights/train_log_rife_426_heavy"),
             "/root/reference/weights/train_log_rife_426_heavy"]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "flownet.pkl")):
            return c
    return None


# ---- GMFSS (models/model_gmfss: FeatureNet.py:6-33, MetricNet.py:23-43, FusionNet.py:6-105) ------------------
def _prelu_conv_pairs(prefix, cin, cout, transposed=False):
    """ResidualBlock / DownsampleBlock / UpsampleBlock parameter shapes (FusionNet.py:6-33)."""
    first = (cin, cout, 4, 4) if transposed else (cout, cin, 3, 3)
    return [(f"{prefix}.0.weight", (1,)), (f"{prefix}.1.weight", first), (f"{prefix}.1.bias", (cout,)),
            (f"{prefix}.2.weight", (1,)), (f"{prefix}.3.weight", (cout, cout, 3, 3)), (f"{prefix}.3.bias", (cout,))]


def gmfss_param_shapes(union=False):
    """union=True: models/model_gmfss_union (image head `residual_model_head0` with 9 input channels)."""
    feat, metric, fusion = [], [], []
    for name, cin, c in (("block1", 3, 64), ("block2", 64, 128), ("block3", 128, 192)):
        feat += _prelu_conv_pairs(name, cin, c)
    metric += [("metric_in.weight", (64, 14, 3, 3)), ("metric_in.bias", (64,))]
    for i in (1, 2, 3):
        metric += [(f"metric_net{i}.0.weight", (1,)), (f"metric_net{i}.1.weight", (64, 64, 3, 3)), (f"metric_net{i}.1.bias", (64,))]
    metric += [("metric_out.0.weight", (1,)), ("metric_out.1.weight", (2, 64, 3, 3)), ("metric_out.1.bias", (2,))]
    for name, cin, c in ((("head0", 9, 64) if union else ("head", 12, 64)), ("head1", 128, 64), ("head2", 256, 128), ("head3", 384, 192),
                         ("01", 64, 64), ("04", 64, 64), ("05", 64, 64), ("11", 128, 128), ("14", 128, 128), ("15", 128, 128),
                         ("21", 192, 192), ("24", 192, 192), ("25", 192, 192)):
        fusion += _prelu_conv_pairs("residual_model_" + name, cin, c)
    fusion += [("residual_model_tail.conv_before_upsample.0.weight", (64, 64, 3, 3)), ("residual_model_tail.conv_before_upsample.0.bias", (64,)),
               ("residual_model_tail.conv_before_upsample.1.weight", (1,)),
               ("residual_model_tail.upsample.0.weight", (256, 64, 3, 3)), ("residual_model_tail.upsample.0.bias", (256,)),
               ("residual_model_tail.conv_last.weight", (3, 64, 3, 3)), ("residual_model_tail.conv_last.bias", (3,))]
    for name, cin, c in (("10", 64, 128), ("20", 128, 192), ("11", 64, 128), ("21", 128, 192)):
        fusion += _prelu_conv_pairs("downsample_model_" + name, cin, c)
    for name, cin, c in (("04", 128, 64), ("14", 192, 128), ("05", 128, 64), ("15", 192, 128)):
        fusion += _prelu_conv_pairs("upsample_model_" + name, cin, c, transposed=True)
    return {"feat": feat, "metric": metric, "fusionnet": fusion}


def synth_gmfss_state(seed=0, union=False):
    """Seeded stand-in weights of the GMFSS nets (same names / shapes as feat.pkl, metric.pkl, fusionnet.pkl)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(2000 + int(seed) + (500 if union else 0))
    out = {}
    for net, shapes in gmfss_param_shapes(union).items():
        sd = {}
        for name, shape in shapes:
            if shape == (1,):
                v = 0.1 + 0.3 * torch.rand(shape, generator=g)              # PReLU slope
            elif name.endswith("bias"):
                v = 0.05 * torch.randn(shape, generator=g)
            else:
                transposed = len(shape) == 4 and shape[2] == 4
                fan_in = (shape[0] * 4) if transposed else shape[1] * shape[2] * shape[3]
                v = torch.randn(shape, generator=g) * (1.2 / fan_in) ** 0.5
            sd[name] = v.float()
        out[net] = sd
    return out


def load_gmfss_state(weights_dir):
    """feat.pkl / metric.pkl / fusionnet.pkl / flownet.pkl (GMFlow) as models/model_gmfss/GMFSS.py:44-56 loads them."""
    out = {}
    for net in ("feat", "metric", "fusionnet", "flownet"):
        path = os.path.join(weights_dir, net + ".pkl")
        if net == "flownet" and not os.path.isfile(path):
            continue
        raw = torch.load(path, map_location="cpu")
        out[net] = {k: v.detach().float().contiguous() for k, v in raw.items()}
    return out


def find_gmfss_weights(explicit=None):
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cands = [explicit, os.environ.get("DRBA_GMFSS_WEIGHTS"), "weights/train_log_gmfss",
             os.path.join(here, "weights/train_log_gmfss"), os.path.join(here, "baseline/_ref/weights/train_log_gmfss"),
             "/root/reference/weights/train_log_gmfss"]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "fusionnet.pkl")):
            return c
    return None


def load_gmflow_state(weights_dir):
    raw = torch.load(os.path.join(weights_dir, "flownet.pkl"), map_location="cpu")
    return {k: v.detach().float().contiguous() for k, v in raw.items()}


def find_gmflow_weights(explicit=None):
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cands = [explicit, os.environ.get("DRBA_GMFSS_WEIGHTS"), "weights/train_log_gmfss",
             os.path.join(here, "weights/train_log_gmfss"), os.path.join(here, "baseline/_ref/weights/train_log_gmfss"),
             "/root/reference/weights/train_log_gmfss"]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "flownet.pkl")):
            return c
    return None
