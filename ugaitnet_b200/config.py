"""Static description of a UGaitNet model (mirror of the reference builder arguments)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

MERGE_MAX, MERGE_AVG, MERGE_SIGNMAX = 0, 1, 2
ACT_LINEAR, ACT_RELU, ACT_LEAKY = 0, 1, 2

BRANCH_NAMES = ("ofBranch", "grayBranch", "depthBranch")


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


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

    @property
    def flat(self) -> int:
        last = self.layers(0, 1)[-1]
        return last["co"] * last["hp"] * last["hp"]
