"""Box-adaptive heads of the similarity branch -- mirror of hdn/models/head/ban.py (DepthwiseXCorr :51-78,
DepthwiseBAN :81-90, MultiBAN :92-127).  UPChannelBAN is not mirrored (no shipped configuration uses it).

Same module tree / state-dict keys (box{2,3,4}.{cls,loc}.{conv_kernel,conv_search,head}.*, cls_weight, loc_weight,
loc_scale) and the same arithmetic, restructured for the GPU:
  * the template side `conv_kernel(z_f)` depends only on the template, which is fixed between `template()` calls;
    the reference recomputes it every frame (ban.py:74).  `MultiBAN.prepare(z_fs)` computes the 6 kernels once;
  * the 3 levels x {cls, loc} correlations have one shape, so they go out as ONE `xcorr_depthwise_multi` launch
    (hdn_xcorr_dw_multi_f32) instead of 6 grouped-conv calls.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from hdn.core.xcorr import xcorr_depthwise, xcorr_depthwise_multi
from hdn_b200 import convs, ops
from hdn_b200.convs import conv_bn_act


class BAN(nn.Module):
    def forward(self, z_f, x_f):
        raise NotImplementedError


def _conv_bn_relu(cin, cout, k):
    return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=k, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class DepthwiseXCorr(nn.Module):
    correlate = staticmethod(xcorr_depthwise)
    circular = False

    def __init__(self, in_channels, hidden, out_channels, kernel_size=3):
        super().__init__()
        self.conv_kernel = _conv_bn_relu(in_channels, hidden, kernel_size)
        self.conv_search = _conv_bn_relu(in_channels, hidden, kernel_size)
        self.head = nn.Sequential(nn.Conv2d(hidden, hidden, kernel_size=1, bias=False), nn.BatchNorm2d(hidden), nn.ReLU(inplace=True),
                                  nn.Conv2d(hidden, out_channels, kernel_size=1))

    # conv + BN + ReLU of each Sequential as one fused tensor-core launch where eligible (hdn_b200.convs.conv_bn_act)
    def kernel_features(self, z):
        return conv_bn_act(self.conv_kernel[0], self.conv_kernel[1], z, relu=True)

    def search_features(self, x):
        return conv_bn_act(self.conv_search[0], self.conv_search[1], x, relu=True)

    def predict(self, feature):
        return self.head[3](conv_bn_act(self.head[0], self.head[1], feature, relu=True))

    def forward(self, kernel, search):
        return self.predict(self.correlate(self.search_features(search), self.kernel_features(kernel)))


class DepthwiseBAN(BAN):
    branch = DepthwiseXCorr
    loc_channels = 2

    def __init__(self, in_channels=256, out_channels=256, cls_out_channels=2, weighted=False):
        super().__init__()
        self.cls = self.branch(in_channels, out_channels, cls_out_channels)
        self.loc = self.branch(in_channels, out_channels, self.loc_channels)

    def forward(self, z_f, x_f):
        return self.cls(z_f, x_f), self.loc(z_f, x_f)


class MultiBAN(BAN):
    level_head = DepthwiseBAN

    def __init__(self, in_channels, cls_out_channels, weighted=False):
        super().__init__()
        self.weighted = weighted
        self.levels = len(in_channels)
        for i, ch in enumerate(in_channels):
            self.add_module("box%d" % (i + 2), self.level_head(ch, ch, cls_out_channels))
        if weighted:
            self.cls_weight = nn.Parameter(torch.ones(self.levels))
            self.loc_weight = nn.Parameter(torch.ones(self.levels))
        self.loc_scale = nn.Parameter(torch.ones(self.levels))
        self._kernels = None

    def _branches(self):
        for i in range(self.levels):
            box = getattr(self, "box%d" % (i + 2))
            yield box.cls
            yield box.loc

    def prepare(self, z_fs):
        """Template-side kernels, once per template: [cls2, loc2, cls3, loc3, cls4, loc4]."""
        self._kernels = [br.kernel_features(z_fs[n // 2]).contiguous() for n, br in enumerate(self._branches())]
        return self._kernels

    # ------------------------------------------------------------------ fused path
    # conv_search x6 (one tcgen05 launch) -> correlation x6 (one launch) -> head 1x1 + BN + ReLU + 1x1 x6 (one launch, the hidden
    # 256-channel map stays on chip) -> level-weighted sum (+ K6 arg-max) (one launch): 4 launches instead of ~40 per stage.
    def _fused_state(self):
        """Packed weights / folded BatchNorms of the search-side layers, the second 1x1 convolutions and the level weights;
        re-derived when a parameter changes (held on the modules, like convs._folded)."""
        branches = list(self._branches())
        ver = tuple((p.data_ptr(), p._version) for p in self.parameters()) + tuple((b.data_ptr(), b._version) for b in self.buffers())
        hit = self.__dict__.get("_hdn_fused")
        if hit is not None and hit[0] == ver:
            return hit[1]
        search = [convs._folded(br.conv_search[0], br.conv_search[1]) for br in branches]
        hidden = [convs._folded(br.head[0], br.head[1]) for br in branches]
        st = {
            "search": [list(col) for col in zip(*search)], "hidden": [list(col) for col in zip(*hidden)],
            "w2": [br.head[3].weight.detach().reshape(br.head[3].out_channels, -1).float().contiguous() for br in branches],
            "b2": [br.head[3].bias.detach().float().contiguous() for br in branches],
            "loc_scale": [float(v) for v in self.loc_scale.detach().cpu()],
        }
        if self.weighted:  # ban.py:112-125
            st["cls_w"] = [float(v) for v in F.softmax(self.cls_weight.detach(), 0).cpu()]
            st["loc_w"] = [float(v) for v in F.softmax(self.loc_weight.detach(), 0).cpu()]
        else:  # plain average over the levels
            st["cls_w"] = st["loc_w"] = [1.0 / self.levels] * self.levels
        self.__dict__["_hdn_fused"] = (ver, st)
        return st

    def fused_eligible(self, x_fs):
        if torch.is_grad_enabled() or self.training or not convs.USE_TENSOR_CORES or 2 * self.levels > 8 or self.levels > 4:
            return False
        branches = list(self._branches())
        x0 = x_fs[0]
        ok = all(x.is_cuda and x.dtype == torch.float32 and tuple(x.shape) == tuple(x0.shape) for x in x_fs)
        for br in branches:
            ok = ok and convs.tensor_core_eligible(br.conv_search[0], x0) and br.head[0].in_channels % 128 == 0 \
                and br.head[0].in_channels == br.head[0].out_channels == br.conv_search[0].out_channels and br.head[3].bias is not None
        return bool(ok)

    def out_size(self, x_fs, kernels):
        """Side of the score map: search features -> 3x3 valid conv -> (circular) correlation with the template kernels."""
        return ops.xcorr_out_hw(x_fs[0].shape[-2] - 2, x_fs[0].shape[-1] - 2, kernels[0].shape[-2], kernels[0].shape[-1], self._circular())[0]

    def _circular(self):
        return next(self._branches()).circular

    def fused(self, x_fs, kernels, window=None, win_influence=0.0, want_maps=True):
        """-> (cls, loc, packed K6 buffer).  kernels: the template-side correlation kernels of `prepare`."""
        st = self._fused_state()
        branches = list(self._branches())
        xs = [x_fs[n // 2] for n in range(len(branches))]
        searches = ops.conv_gemm_multi(xs, *st["search"], ksize=3, dilation=branches[0].conv_search[0].dilation[0], relu=True, valid=True)
        feats = xcorr_depthwise_multi(searches, kernels, circular=branches[0].circular)
        if len({tuple(w.shape) for w in st["w2"]}) == 1:  # cls and loc have the same width (similarity head: 2 and 2): one launch
            parts = ops.head_project_multi(feats, st["hidden"][0], st["hidden"][1], st["hidden"][2], st["w2"])
            n_cls, n_loc = parts[0::2], parts[1::2]
        else:
            n_cls = ops.head_project_multi(feats[0::2], st["hidden"][0][0::2], st["hidden"][1][0::2], st["hidden"][2][0::2], st["w2"][0::2])
            n_loc = ops.head_project_multi(feats[1::2], st["hidden"][0][1::2], st["hidden"][1][1::2], st["hidden"][2][1::2], st["w2"][1::2])
        N = feats[0].shape[-1]
        return ops.head_score(n_cls, n_loc, st["b2"][0::2], st["b2"][1::2], st["cls_w"], st["loc_scale"], st["loc_w"], N, window, win_influence,
                              want_maps)

    def forward(self, z_fs, x_fs, kernels=None):
        if kernels is None:
            kernels = [br.kernel_features(z_fs[n // 2]) for n, br in enumerate(self._branches())]
        if self.fused_eligible(x_fs) and len({tuple(k.shape) for k in kernels}) == 1 and x_fs[0].shape[-1] == x_fs[0].shape[-2]:
            cls, loc, _ = self.fused(x_fs, kernels)
            return cls, loc
        branches = list(self._branches())
        searches = [br.search_features(x_fs[n // 2]) for n, br in enumerate(branches)]
        same = len({tuple(s.shape) for s in searches}) == 1 and len({tuple(k.shape) for k in kernels}) == 1
        if same and len(searches) <= 8:
            feats = xcorr_depthwise_multi(searches, kernels, circular=branches[0].circular)
        else:
            feats = [br.correlate(s, k) for br, s, k in zip(branches, searches, kernels)]
        outs = [br.predict(f) for br, f in zip(branches, feats)]
        cls = outs[0::2]
        loc = [l * self.loc_scale[i] for i, l in enumerate(outs[1::2])]  # ban.py:109
        if self.weighted:  # ban.py:112-125
            cw, lw = F.softmax(self.cls_weight, 0), F.softmax(self.loc_weight, 0)
            return sum(c * cw[i] for i, c in enumerate(cls)), sum(l * lw[i] for i, l in enumerate(loc))
        return sum(cls) / len(cls), sum(loc) / len(loc)
