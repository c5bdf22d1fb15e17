"""Pixel-grid layouts of the NHWC bf16 activations and the tap tables that turn every convolution of the
path into a shifted-row GEMM.

A *grid* buffer is a 2-D array ``[planes * B * Hg * Wg, ld]``. A logical pixel ``(b, h, w)`` of the
``H x W`` content lives at grid position ``(h + h0, w + w0)``; ``h`` may range over ``[-pad_lo, H + pad_hi)``
when the grid carries a halo (reflect-mirrored or zero). With ``phase=True`` the grid is split into four
parity planes (row parity, column parity) so that a stride-2 convolution reads unit-stride rows.
"""
from dataclasses import dataclass, field
from typing import List, Tuple


@dataclass(frozen=True)
class Lay:
    B: int
    H: int          # logical content height
    W: int
    Hg: int         # grid rows per image (per plane when phase)
    Wg: int
    h0: int = 0     # content offset inside the (un-split) grid
    w0: int = 0
    phase: bool = False
    ld: int = 0     # channels per row (row pitch in elements)
    c0: int = 0     # channel offset of this view inside the row
    C: int = 0      # channels of this view

    @property
    def plane_rows(self) -> int:
        return self.B * self.Hg * self.Wg

    @property
    def rows(self) -> int:
        return self.plane_rows * (4 if self.phase else 1)

    def view(self, c0: int, C: int) -> "Lay":
        return Lay(self.B, self.H, self.W, self.Hg, self.Wg, self.h0, self.w0, self.phase, self.ld,
                   self.c0 + c0, C)

    def row(self, b: int, h: int, w: int) -> int:
        hp, wp = h + self.h0, w + self.w0
        if not self.phase:
            return (b * self.Hg + hp) * self.Wg + wp
        return ((hp & 1) * 2 + (wp & 1)) * self.plane_rows + (b * self.Hg + (hp >> 1)) * self.Wg + (wp >> 1)


def pad_to(c: int, mult: int = 16) -> int:
    return (c + mult - 1) // mult * mult


def chan_pad(c: int) -> int:
    """Channel count the tensor-core kernels accept: 16, 32, 48 or a multiple of 64."""
    p = pad_to(c, 16)
    if p <= 48:
        return p
    return pad_to(c, 64)


# ---------------------------------------------------------------------------------------------------
# Convolution geometries.  Every entry gives, for a logical input of H x W:
#   in_lay(B, C)   layout the producer must write (the conv's A operand)
#   out_lay(B, N)  layout of the raw conv output == layout of its gradient dY
#   fwd taps       [(tap slot, row shift)] per launch (one launch, or four for a transposed conv)
#   bwd taps       same for the data gradient (A operand = dY, output = gradient of in_lay)
# tap slot t = kh * k + kw indexes the packed weights [T][N][C].
# ---------------------------------------------------------------------------------------------------
@dataclass
class Launch:
    taps: List[Tuple[int, int]]        # (tap slot, row shift)
    out_plane: int = 0                 # output plane (phase layouts) this launch writes


@dataclass
class ConvGeom:
    kind: str            # 's1', 's2', 'up'
    k: int
    pad: int
    pad_mode: str        # 'reflect' | 'zero'
    B: int
    H: int               # logical input size
    W: int
    Ho: int
    Wo: int
    in_lay: Lay = None
    out_lay: Lay = None
    fwd: List[Launch] = field(default_factory=list)
    bwd: List[Launch] = field(default_factory=list)
    in_pad_lo: int = 0
    in_pad_hi: int = 0


def geom_s1(B, H, W, k, pad_mode, Cin_ld, Cout_ld) -> ConvGeom:
    """k x k stride-1 'same' convolution with padding (k-1)/2 (reflect or zero)."""
    p = (k - 1) // 2
    Hg, Wg = H + 2 * p, W + 2 * p
    g = ConvGeom('s1', k, p, pad_mode, B, H, W, H, W, in_pad_lo=p, in_pad_hi=p)
    g.in_lay = Lay(B, H, W, Hg, Wg, p, p, False, Cin_ld, 0, Cin_ld)
    g.out_lay = Lay(B, H, W, Hg, Wg, 0, 0, False, Cout_ld, 0, Cout_ld)
    g.fwd = [Launch([(kh * k + kw, kh * Wg + kw) for kh in range(k) for kw in range(k)])]
    g.bwd = [Launch([(kh * k + kw, -(kh * Wg + kw)) for kh in range(k) for kw in range(k)])]
    return g


def geom_s2(B, H, W, Cin_ld, Cout_ld) -> ConvGeom:
    """3 x 3 stride-2 zero-pad-1 convolution (H, W even): input in four parity planes."""
    assert H % 2 == 0 and W % 2 == 0
    Hg, Wg = H // 2 + 1, W // 2 + 1
    g = ConvGeom('s2', 3, 1, 'zero', B, H, W, H // 2, W // 2, in_pad_lo=1, in_pad_hi=1)
    g.in_lay = Lay(B, H, W, Hg, Wg, 1, 1, True, Cin_ld, 0, Cin_ld)
    g.out_lay = Lay(B, H // 2, W // 2, Hg, Wg, 0, 0, False, Cout_ld, 0, Cout_ld)
    pr = B * Hg * Wg
    g.fwd = [Launch([(kh * 3 + kw, ((kh & 1) * 2 + (kw & 1)) * pr + (kh >> 1) * Wg + (kw >> 1))
                     for kh in range(3) for kw in range(3)])]
    g.bwd = []
    for ph in range(2):
        for pw in range(2):
            taps = [(kh * 3 + kw, -((kh >> 1) * Wg + (kw >> 1)))
                    for kh in range(3) for kw in range(3) if (kh & 1) == ph and (kw & 1) == pw]
            g.bwd.append(Launch(taps, out_plane=ph * 2 + pw))
    return g


def geom_up(B, H, W, Cin_ld, Cout_ld) -> ConvGeom:
    """ConvTranspose2d(k=3, stride=2, padding=1, output_padding=1): output 2H x 2W in four parity planes."""
    Hg, Wg = H + 1, W + 1
    g = ConvGeom('up', 3, 1, 'zero', B, H, W, 2 * H, 2 * W, in_pad_lo=0, in_pad_hi=1)
    g.in_lay = Lay(B, H, W, Hg, Wg, 0, 0, False, Cin_ld, 0, Cin_ld)
    g.out_lay = Lay(B, 2 * H, 2 * W, Hg, Wg, 0, 0, True, Cout_ld, 0, Cout_ld)
    pr = B * Hg * Wg
    g.fwd = []
    for py in range(2):
        for px in range(2):
            kys = [1] if py == 0 else [0, 2]
            kxs = [1] if px == 0 else [0, 2]
            taps = [(ky * 3 + kx, (1 if ky == 0 else 0) * Wg + (1 if kx == 0 else 0)) for ky in kys for kx in kxs]
            g.fwd.append(Launch(taps, out_plane=py * 2 + px))
    taps = []
    for ky in range(3):
        for kx in range(3):
            py, px = (0 if ky == 1 else 1), (0 if kx == 1 else 1)
            taps.append((ky * 3 + kx, (py * 2 + px) * pr - (1 if ky == 0 else 0) * Wg - (1 if kx == 0 else 0)))
    g.bwd = [Launch(taps)]
    return g
