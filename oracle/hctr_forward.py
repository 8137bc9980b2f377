"""oracle/hctr_forward.py - TEST INFRASTRUCTURE, NOT PRODUCT CODE.

torch fp32 functional restatement of the reference model's forward from a reference-layout state_dict
(models/handwritten_ctr_model.py). Floating-point kernels are checked against this within the tolerances
stated in the tests. Pinned against the imported reference module by tests/golden (make_golden.py)."""
import torch
import torch.nn.functional as F

STAGE_BLOCKS = (2, 4, 5, 1)          # hctr_model.__init__ :166
EPS = 1e-5                           # nn.BatchNorm2d default


def _bn(x, sd, name, train_stats=None):
    if train_stats is not None:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        train_stats[name] = (mean, x.var(dim=(0, 2, 3), unbiased=True))
    else:
        # eval mode: the same fused op nn.BatchNorm2d dispatches to (keeps the CPU baseline as fast as the reference)
        return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"], sd[name + ".bias"],
                            False, 0.0, EPS)
    return ((x - mean.view(1, -1, 1, 1)) * torch.rsqrt(var.view(1, -1, 1, 1) + EPS)
            * sd[name + ".weight"].view(1, -1, 1, 1) + sd[name + ".bias"].view(1, -1, 1, 1))


def _conv(x, sd, name, pad):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), padding=pad)


def _se(x, sd, name):                                                  # SELayer.forward :26-30
    y = x.mean(dim=(2, 3))
    y = torch.relu(y @ sd[name + ".fc.0.weight"].t())
    y = torch.sigmoid(y @ sd[name + ".fc.2.weight"].t())
    return x * y.view(y.shape[0], y.shape[1], 1, 1)


def _relu(x):
    # nn.ReLU(inplace=True) in the reference (:42,94); in place only when autograd is not recording
    return torch.relu_(x) if not torch.is_grad_enabled() else torch.relu(x)


def _block(x, sd, name, train_stats):                                   # BasicBlock.forward :47-60 (dropout = identity)
    out = _relu(_bn(_conv(x, sd, name + ".conv1", 1), sd, name + ".bn1", train_stats))
    out = _bn(_conv(out, sd, name + ".conv2", 1), sd, name + ".bn2", train_stats)
    out = _se(out, sd, name + ".se")
    if (name + ".downsample.0.weight") in sd:
        res = _bn(_conv(x, sd, name + ".downsample.0", 0), sd, name + ".downsample.1", train_stats)
    else:
        res = x
    return _relu(out.add_(res) if not torch.is_grad_enabled() else out + res)


def features(x, sd, train_stats=None, taps=None):
    """ResNet.forward :115-153 in eval mode (or with batch statistics when train_stats is a dict)."""
    def tap(name, v):
        if taps is not None:
            taps[name] = v
        return v
    x = tap("bn0_1", _relu(_bn(_conv(x, sd, "cnn.conv0_1", 1), sd, "cnn.bn0_1", train_stats)))
    x = _relu(_bn(_conv(x, sd, "cnn.conv0_2", 1), sd, "cnn.bn0_2", train_stats))
    x = tap("pool0", F.max_pool2d(x, (2, 1), (2, 1)))
    for s, n in enumerate(STAGE_BLOCKS, start=1):
        for i in range(n):
            x = tap("block%d.%d" % (s, i), _block(x, sd, "cnn.block%d.%d" % (s, i), train_stats))
        x = _relu(_bn(_conv(x, sd, "cnn.conv%d" % s, 1), sd, "cnn.bn%d" % s, train_stats))
        x = tap("pool%d" % s, F.max_pool2d(x, (2, 1), (2, 1)))
    return x


def forward(x, sd, train_stats=None, taps=None):
    """hctr_model.forward :171-178 -> [W, B, C] fp32."""
    f = features(x, sd, train_stats, taps)
    f = f.flatten(1, 2).permute(0, 2, 1)                                # d = c*4 + h
    y = f @ sd["linear.weight"].t() + sd["linear.bias"]
    return y.permute(1, 0, 2)


def calibrate_bn(sd, x):
    """BN-calibrated random weights: one train-mode pass with momentum=1.0, dropout off;
    running stats := batch mean / unbiased batch variance. Returns a new state_dict."""
    stats = {}
    with torch.no_grad():
        forward(x, sd, train_stats=stats)
    out = dict(sd)
    for name, (mean, var_unbiased) in stats.items():
        out[name + ".running_mean"] = mean.clone()
        out[name + ".running_var"] = var_unbiased.clone()
    return out


def random_state_dict(num_classes=7375, seed=1234):
    """Random-init parameters of the reference architecture in the reference's state_dict layout,
    PyTorch-default ranges (uniform +-1/sqrt(fan_in); BN gamma=1, beta=0, mean=0, var=1). Used by bench.py's CPU arm so
    that it does not touch the product package at all."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def uni(shape, fan_in):
        b = 1.0 / (fan_in ** 0.5)
        return (torch.rand(shape, generator=g) * 2 - 1) * b

    def conv(name, cout, cin, k, bias=True):
        sd[name + ".weight"] = uni((cout, cin, k, k), cin * k * k)
        if bias:
            sd[name + ".bias"] = uni((cout,), cin * k * k)

    def bn(name, c):
        sd[name + ".weight"] = torch.ones(c); sd[name + ".bias"] = torch.zeros(c)
        sd[name + ".running_mean"] = torch.zeros(c); sd[name + ".running_var"] = torch.ones(c)

    conv("cnn.conv0_1", 64, 1, 3); bn("cnn.bn0_1", 64)
    conv("cnn.conv0_2", 64, 64, 3); bn("cnn.bn0_2", 64)
    width = 64
    for s, (planes, n) in enumerate(zip((128, 256, 512, 512), STAGE_BLOCKS), start=1):
        for i in range(n):
            name = "cnn.block%d.%d" % (s, i)
            cin = width if i == 0 else planes
            conv(name + ".conv1", planes, cin, 3); bn(name + ".bn1", planes)
            conv(name + ".conv2", planes, planes, 3); bn(name + ".bn2", planes)
            sd[name + ".se.fc.0.weight"] = uni((planes // 16, planes), planes)
            sd[name + ".se.fc.2.weight"] = uni((planes, planes // 16), planes // 16)
            if i == 0 and width != planes:
                conv(name + ".downsample.0", planes, width, 1, bias=False); bn(name + ".downsample.1", planes)
        width = planes
        conv("cnn.conv%d" % s, planes, planes, 3); bn("cnn.bn%d" % s, planes)
    sd["linear.weight"] = uni((num_classes, 2048), 2048)
    sd["linear.bias"] = uni((num_classes,), 2048)
    return sd
