"""Static description of a UGaitNet model (mirror of the reference builder arguments)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

MERGE_MAX, MERGE_AVG, MERGE_SIGNMAX = 0, 1, 2
FUSE3_NO_NORM = 0x100            # include/ugaitnet_b200.h: UGN_FUSE3_NO_NORM (gate + fusion without l2_normalize)
ACT_LINEAR, ACT_RELU, ACT_LEAKY = 0, 1, 2

BRANCH_NAMES = ("ofBranch", "grayBranch", "depthBranch")
AUX_NAMES = ("classprob_of", "classprob_gray", "classprob_depth")


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# build_3Dbranch (nets/mj_uwyhNets_ba.py:346-363): (kernel (kt,kh,kw), strides) of the six activated Conv3D layers
LAYERS3D = (((3, 5, 5), (1, 2, 2)), ((3, 3, 3), (1, 2, 2)), ((3, 3, 3), (2, 2, 2)), ((3, 3, 3), (2, 2, 2)),
            ((3, 2, 2), (1, 1, 1)), ((2, 1, 1), (1, 1, 1)))


@dataclass
class NetConfig:
    """Arguments of UWYHSemiNet3Mods.build / UWYHSemiNet.build that shape the graph
    (/root/reference/nets/mj_uwyhNets_ba.py:1032-1037, :669-672)."""
    in_channels: Sequence[int] = (50, 25, 25)
    filters_numbers: Sequence[int] = (96, 192, 512, 512)
    filters_size: Sequence[int] = (7, 5, 3, 2)
    nd: int = 2048                 # ndense_units[0] (signature dimension)
    nc: int = 0                    # ndense_units[1] (FC1 "code"), 0 = absent
    nclasses: int = 150
    weight_decay: float = 5e-5
    merge: int = MERGE_MAX         # Keras default fMerge=Maximum
    act: int = ACT_RELU
    alpha: float = 0.3
    margin: float = 0.2
    wver: float = 1.0              # loss_weights[0]
    wid: float = 1.0               # loss_weights[1]
    hw: int = 60
    dropout: float = 0.0
    single: bool = False           # 1-modality graph: no gate / fusion / l2_normalize (:900-915)
    label_smoothing: float = 0.0   # smoothlabels -> tf.losses.CategoricalCrossentropy(label_smoothing) (:1252-1262)
    normbfmerge: bool = False      # l2_normalize every branch output before its gate (:1167-1168)
    aux_losses: bool = False       # extra Dense(nclasses, softmax) head + CE on every gated branch output (:1222-1251)
    waux: float = 1.0              # their loss weight: loss_weights[-1] (:1264-1268)
    triplet_hard: bool = False     # compile_hard (:1302-1306): tfa TripletHardLoss (batch-hard) instead of batch-all
    branch3d: Sequence[bool] = ()  # use3D (:336-417, :1077-1099): per modality True -> Conv3D branch on [B,25,60,60,1] (frames =
    #                                in_channels[m]); fp32 validation engine only
    filters3d: Sequence[int] = (64, 128, 256, 512, 512, 512)    # its six activated Conv3D layers (geometry: LAYERS3D)
    pair_loss: bool = False        # UWYHNet.build (:154-245): rows [0,B) / [B,2B) = the two sides of B pairs, labels [B] in
    #                                {0,1}, VerifLossLayer (nets/mj_loss.py:65-95) instead of the triplet loss; nclasses == 0
    postriplet: int = 1            # 2 (needs nc > 0; 2-modality builder, :814-832): the fusion is NOT normalised, FC1 is the
    #                                layer "signature", its l2_normalize ("code") is what the triplet loss and the classifier see

    @property
    def nmods(self) -> int:
        return len(self.in_channels)

    def layers(self, m: int, pad: int) -> List[dict]:
        """Geometry of the conv stack of modality m; channel counts padded to `pad`."""
        out = []
        s = self.hw
        cin = self.in_channels[m]
        n = len(self.filters_numbers)
        for i, (co, k) in enumerate(zip(self.filters_numbers, self.filters_size)):
            ho = s - k + 1
            pool = i != n - 1
            hp = ho // 2 if pool else ho
            out.append(dict(cin=cin, cp=round_up(cin, pad) if pad > 1 else cin, h=s, k=k, co=co, ho=ho,
                            hp=hp, pool=pool))
            s, cin = hp, co
        return out

    def is3d(self, m: int) -> bool:
        return bool(len(self.branch3d) > m and self.branch3d[m])

    def layers3d(self, m: int) -> List[dict]:
        """Geometry of the Conv3D stack of modality m (build_3Dbranch, nets/mj_uwyhNets_ba.py:346-363): strided 'valid'
        channels-last convolutions that take the [25,60,60,1] volume down to 1x1x1."""
        out = []
        t, h, cin = self.in_channels[m], self.hw, 1
        for co, (k, s) in zip(self.filters3d, LAYERS3D):
            to, ho = (t - k[0]) // s[0] + 1, (h - k[1]) // s[1] + 1
            out.append(dict(cin=cin, co=co, k=k, s=s, t=t, h=h, to=to, ho=ho))
            t, h, cin = to, ho, co
        if (t, h) != (1, 1):
            raise ValueError(f"use3D expects {self.in_channels[m]} x {self.hw} x {self.hw} volumes that end at 1x1x1, got {t}x{h}x{h}")
        return out

    @property
    def flat(self) -> int:
        last = self.layers(0, 1)[-1]
        return last["co"] * last["hp"] * last["hp"]


# (name, cin, cout, k) of build_gaitset_branch, in graph order
# (/root/reference/nets/mj_uwyhNets_ba.py:427-466); cin None = the per-frame input channels (1 | 2)
GS_CONVS = (("a1", None, 32, 5), ("a2", 32, 32, 3), ("b1", 32, 64, 3), ("b2", 64, 64, 3), ("a3", 32, 64, 3),
            ("a4", 64, 64, 3), ("b3", 64, 128, 3), ("b4", 128, 128, 3), ("a5", 64, 128, 3), ("a6", 128, 128, 3))
GS_PARTS = 62                      # 2 * (1 + 2 + 4 + 8 + 16) HPP strips (:468-479)
GS_ALPHA = 0.3                     # layers.LeakyReLU() default


@dataclass
class GaitSetConfig:
    """Arguments of UWYHSemiNet3Mods.build(..., gaitset=True) that shape the graph
    (/root/reference/nets/mj_uwyhNets_ba.py:1032-1037; branch :420-484)."""
    in_channels: Sequence[int] = (2, 1, 1)    # per-frame channels: OF (x,y), gray, depth | silhouette
    frames: int = 25
    hw: int = 60
    hidden: int = 256              # MatMul hidden_dim (:24)
    nc: int = 0                    # ndense_units[1] (FC1 "code"), 0 = absent
    nclasses: int = 150
    merge: int = MERGE_MAX
    alpha: float = 0.3             # LeakyReLU after "code" (gaitset needs fActivation != 'relu', :1101)
    margin: float = 0.2
    wver: float = 1.0
    wid: float = 1.0
    dropout: float = 0.0           # Dropout(name="dropcode") after FC1 (:1203); the branches have none
    single: bool = False           # UWYHSemiNet.build on ONE shape (:890-905): the branch output is the signature
    label_smoothing: float = 0.0   # smoothlabels (:1252-1262)
    postriplet: int = 1            # 2 (needs nc > 0; 2-modality builder :814-832): un-normalised fusion -> Dense "signature"
    #                                -> LeakyReLU -> l2_normalize(axis=1) "code" = the embedding of the triplet loss / classifier

    @property
    def nmods(self) -> int:
        return len(self.in_channels)

    @property
    def nd(self) -> int:           # signature width per part (what the stacked-CNN config calls nd)
        return self.hidden
