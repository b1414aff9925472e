"""Minimal protobuf-text parser for the two Bayesian SegNet prototxts.

Host-side mirror of the topology that `caffe::Net<float>(model_file, TEST)` builds in
the reference (`src/bayesian_segnet/bayesian_segnet.cpp:59-61`).  Only the nine layer
types the shipped prototxts use are understood (`config/bayesian_segnet/*/kitti/*.prototxt`):
Convolution, ReLU, BN (INFERENCE), LRN, Pooling (MAX 2x2/2 + mask), Upsample, Dropout,
Softmax.  The C++ twin of this parser lives in `sivo_b200/csrc/prototxt.cc`; this one is
used by the oracle, the synthetic-weight generator and the tests.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

_TOKEN = re.compile(r'\s*(?:(#[^\n]*)|([{}:])|"((?:[^"\\]|\\.)*)"|([^\s{}:#"]+))')


def _tokens(text: str):
    pos = 0
    n = len(text)
    while pos < n:
        m = _TOKEN.match(text, pos)
        if not m:
            if text[pos:].strip() == "":
                return
            raise ValueError(f"prototxt: cannot tokenise at offset {pos}: {text[pos:pos+30]!r}")
        pos = m.end()
        if m.group(1) is not None:
            yield ("nl", None)  # a comment ends the current scalar (blank `dim:` case)
            continue
        if m.group(2) is not None:
            yield ("p", m.group(2))
        elif m.group(3) is not None:
            yield ("s", m.group(3))
        else:
            yield ("w", m.group(4))


class Message(dict):
    """field name -> list of values (scalars as str, sub-messages as Message)."""

    def one(self, key: str, default: Any = None) -> Any:
        v = self.get(key)
        return v[0] if v else default

    def all(self, key: str) -> List[Any]:
        return self.get(key, [])


def parse(text: str) -> Message:
    toks = [t for t in _tokens(text)]
    pos = 0

    def parse_msg(depth: int) -> Message:
        nonlocal pos
        msg = Message()
        while pos < len(toks):
            kind, val = toks[pos]
            if kind == "nl":
                pos += 1
                continue
            if kind == "p" and val == "}":
                if depth == 0:
                    raise ValueError("prototxt: unbalanced '}'")
                pos += 1
                return msg
            if kind != "w":
                raise ValueError(f"prototxt: expected field name, got {val!r}")
            name = val
            pos += 1
            kind, val = toks[pos]
            if kind == "p" and val == "{":
                pos += 1
                msg.setdefault(name, []).append(parse_msg(depth + 1))
                continue
            if not (kind == "p" and val == ":"):
                raise ValueError(f"prototxt: expected ':' or '{{' after {name}")
            pos += 1
            kind, val = toks[pos]
            if kind == "p" and val == "{":
                pos += 1
                msg.setdefault(name, []).append(parse_msg(depth + 1))
            elif kind in ("s", "w"):
                pos += 1
                msg.setdefault(name, []).append(val)
            else:
                # `dim: # SET SAMPLE SIZE HERE` -- the Standard prototxt ships with a blank
                # first dim (config/bayesian_segnet/standard/kitti/bayesian_segnet_kitti.prototxt:4)
                msg.setdefault(name, []).append(None)
        if depth != 0:
            raise ValueError("prototxt: missing '}'")
        return msg

    return parse_msg(0)


@dataclass
class Layer:
    name: str
    type: str
    bottoms: List[str]
    tops: List[str]
    # Convolution
    num_output: int = 0
    kernel: int = 0
    pad: int = 0
    stride: int = 1
    bias_term: bool = True
    weight_filler: str = "constant"
    # LRN
    local_size: int = 5
    alpha: float = 1.0
    beta: float = 0.75
    k: float = 1.0
    # Dropout
    dropout_ratio: float = 0.5
    sample_weights_test: bool = False
    # Upsample
    scale: int = 2
    # ReLU
    negative_slope: float = 0.0


@dataclass
class NetSpec:
    name: str
    input_name: str
    input_dims: List[Optional[int]]  # N(=T), C, H, W ; N may be None (blank in the Standard prototxt)
    layers: List[Layer] = field(default_factory=list)

    @property
    def T(self) -> Optional[int]:
        return self.input_dims[0]


def _b(v: Any, default: bool) -> bool:
    if v is None:
        return default
    return str(v).lower() in ("true", "1")


def load_net(text: str, T: Optional[int] = None) -> NetSpec:
    """Build the layer list; `T` overrides / fills the first input dim (MC sample count)."""
    m = parse(text)
    name = m.one("name", "")
    input_name = m.one("input", "data")
    if m.all("input_dim"):
        dims = [None if d is None else int(d) for d in m.all("input_dim")]
    elif m.all("input_shape"):
        dims = [None if d is None else int(d) for d in m.one("input_shape").all("dim")]
    else:
        raise ValueError("prototxt: no input_dim / input_shape")
    if len(dims) == 3:  # blank first dim swallowed entirely
        dims = [None] + dims
    if len(dims) != 4:
        raise ValueError(f"prototxt: expected 4 input dims, got {dims}")
    if T is not None:
        dims[0] = int(T)
    net = NetSpec(name=name, input_name=input_name, input_dims=dims)
    if m.all("layers"):
        raise ValueError("prototxt: V1LayerParameter ('layers') is not supported")
    for lm in m.all("layer"):
        ly = Layer(name=lm.one("name"), type=lm.one("type"),
                   bottoms=list(lm.all("bottom")), tops=list(lm.all("top")))
        t = ly.type
        if t == "Convolution":
            cp = lm.one("convolution_param")
            ly.num_output = int(cp.one("num_output"))
            ly.kernel = int(cp.one("kernel_size"))
            ly.pad = int(cp.one("pad", 0))
            ly.stride = int(cp.one("stride", 1))
            ly.bias_term = _b(cp.one("bias_term"), True)
            wf = cp.one("weight_filler")
            ly.weight_filler = wf.one("type", "constant") if wf else "constant"
            if ly.stride != 1 or int(cp.one("group", 1)) != 1 or int(cp.one("dilation", 1)) != 1:
                raise ValueError(f"{ly.name}: only stride 1 / group 1 / dilation 1 convolutions are on the path")
        elif t == "LRN":
            lp = lm.one("lrn_param") or Message()
            ly.local_size = int(lp.one("local_size", 5))
            ly.alpha = float(lp.one("alpha", 1.0))
            ly.beta = float(lp.one("beta", 0.75))
            ly.k = float(lp.one("k", 1.0))
            if (lp.one("norm_region", "ACROSS_CHANNELS")) != "ACROSS_CHANNELS":
                raise ValueError("LRN WITHIN_CHANNEL is not on the path")
        elif t == "Pooling":
            pp = lm.one("pooling_param")
            if pp.one("pool", "MAX") != "MAX" or int(pp.one("kernel_size")) != 2 or int(pp.one("stride", 1)) != 2 \
                    or int(pp.one("pad", 0)) != 0:
                raise ValueError(f"{ly.name}: only MAX 2x2 stride 2 pooling is on the path")
            if len(ly.tops) != 2:
                raise ValueError(f"{ly.name}: pooling must emit a mask top")
        elif t == "Upsample":
            up = lm.one("upsample_param") or Message()
            ly.scale = int(up.one("scale", 2))
            if ly.scale != 2 or up.one("upsample_h") or up.one("pad_out_h") or up.one("scale_h"):
                raise ValueError(f"{ly.name}: only scale-2 upsample is on the path")
        elif t == "Dropout":
            dp = lm.one("dropout_param") or Message()
            ly.dropout_ratio = float(dp.one("dropout_ratio", 0.5))
            ly.sample_weights_test = _b(dp.one("sample_weights_test"), False)
        elif t == "BN":
            bp = lm.one("bn_param") or Message()
            if bp.one("bn_mode", "LEARN") != "INFERENCE":
                raise ValueError(f"{ly.name}: BN must be bn_mode INFERENCE at test time")
        elif t == "ReLU":
            rp = lm.one("relu_param")
            ly.negative_slope = float(rp.one("negative_slope", 0.0)) if rp else 0.0
        elif t == "Softmax":
            pass
        else:
            raise ValueError(f"layer type {t!r} is not on the SIVO perception path")
        net.layers.append(ly)
    return net


def blob_shapes(net: NetSpec) -> Dict[str, tuple]:
    """(C, H, W) of every blob, following Caffe's Reshape rules for the layer types above."""
    _, C, H, W = net.input_dims
    shapes = {net.input_name: (C, H, W)}
    for ly in net.layers:
        c, h, w = shapes[ly.bottoms[0]]
        if ly.type == "Convolution":
            ho = h + 2 * ly.pad - ly.kernel + 1
            wo = w + 2 * ly.pad - ly.kernel + 1
            shapes[ly.tops[0]] = (ly.num_output, ho, wo)
        elif ly.type == "Pooling":
            # Caffe: ceil((h - k) / s) + 1 (pooling_layer.cpp Reshape); even dims on this path
            ho = -(-(h - 2) // 2) + 1
            wo = -(-(w - 2) // 2) + 1
            shapes[ly.tops[0]] = (c, ho, wo)
            shapes[ly.tops[1]] = (c, ho, wo)
        elif ly.type == "Upsample":
            shapes[ly.tops[0]] = (c, h * 2, w * 2)
        else:
            shapes[ly.tops[0]] = (c, h, w)
    return shapes


def param_shapes(net: NetSpec) -> Dict[str, list]:
    """Layer name -> list of blob shapes, the way `Net::CopyTrainedLayersFrom` expects them
    (caffe/src/caffe/net.cpp:750-785): conv = [(Cout,Cin,k,k), (Cout,)], BN = [(1,C,1,1)]*2."""
    shapes = blob_shapes(net)
    out: Dict[str, list] = {}
    for ly in net.layers:
        if ly.type == "Convolution":
            cin = shapes[ly.bottoms[0]][0]
            s = [(ly.num_output, cin, ly.kernel, ly.kernel)]
            if ly.bias_term:
                s.append((ly.num_output,))
            out[ly.name] = s
        elif ly.type == "BN":
            c = shapes[ly.bottoms[0]][0]
            out[ly.name] = [(1, c, 1, 1), (1, c, 1, 1)]
    return out
