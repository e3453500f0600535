"""Offline (no GPU) replica of choose_tile_n() of csrc/gemm_tc.cu over the GEMM shapes of one full-size SDXL UNet forward (Be = 8 samples):
which tile width the picker takes, how many 128-row tiles that makes, how full the last wave of the persistent kernel is.  Quantifies what a
stream-K / split-tail scheduler could recover.  python tools/gemm_wave_model.py"""
import math

SMS = 148
TILES = [64, 96, 128, 144, 160, 192, 208, 224, 240, 256]


def choose_tile_n(m_tiles, N, k_blocks, heavy, only32=True):
    best, best_bn = 1e30, 256
    for bn in TILES:
        if only32 and bn % 32:
            continue
        tiles = m_tiles * math.ceil(N / bn)
        waves = math.ceil(tiles / SMS)
        mma = k_blocks * 4.0 * max(bn / 2.0, (4096.0 + 32.0 * bn) / 91.0)
        epi = bn * (14.0 if heavy else 8.0) + 300.0
        cost = waves * (max(mma, epi) + 150.0) + epi
        if cost < best - 1e-9 or (abs(cost - best) <= 1e-9 and bn > best_bn):
            best, best_bn = cost, bn
    return best_bn


def unet_gemms(Be=8):
    """(label, M, N, K, count, heavy_epilogue) of the linear layers and stride-1 convs of one forward"""
    out = []
    for (hw, c, depth, nblocks) in ((64, 640, 2, 5), (32, 1280, 10, 7)):      # transformer blocks: 2+3 at 64^2, 2+1+3... at 32^2 (+mid)
        M = Be * hw * hw
        n = depth * nblocks
        out += [(f"attn qkv {c}", M, 3 * c, c, n, False), (f"attn out {c}", M, c, c, 2 * n, False), (f"xattn q {c}", M, c, c, n, False),
                (f"ff geglu {c}", M, 8 * c, c, n, True), (f"ff down {c}", M, c, 4 * c, n, False), (f"proj in/out {c}", M, c, c, 2 * nblocks, False)]
    convs = [(128, 320, 320, 4 + 1), (128, 960, 320, 1), (128, 640, 320, 2 + 2), (128, 320, 320, 3),                 # level 0 down + up (conv1 / conv2)
             (64, 320, 640, 1), (64, 640, 640, 3 + 1 + 3), (64, 1920, 640, 1), (64, 1280, 640, 1), (64, 960, 640, 1),      # level 1
             (32, 640, 1280, 1), (32, 1280, 1280, 3 + 2 + 2 + 3 + 1), (32, 2560, 1280, 2), (32, 1920, 1280, 1)]          # level 2 + mid
    for hw, cin, cout, n in convs:
        out.append((f"conv3x3 {cin}->{cout} @{hw}", Be * hw * hw, cout, 9 * cin, n, False))
    return out


def main():
    tot_ideal = tot_model = 0.0
    print(f"{'gemm':28s} {'M':>7s} {'N':>6s} {'K':>6s} {'x':>4s} {'BN':>4s} {'tiles':>6s} {'waves':>6s} {'fill':>6s} {'TFLOP':>7s}")
    for label, M, N, K, count, heavy in unet_gemms():
        m_tiles, kb = math.ceil(M / 128), math.ceil(K / 64)
        bn = choose_tile_n(m_tiles, N, kb, heavy)
        tiles = m_tiles * math.ceil(N / bn)
        waves = tiles / SMS
        fill = waves / math.ceil(waves) * (N / (math.ceil(N / bn) * bn))        # last-wave quantisation x padded columns
        fl = 2.0 * M * N * K * count / 1e12
        tot_ideal += fl
        tot_model += fl / fill
        print(f"{label:28s} {M:7d} {N:6d} {K:6d} {count:4d} {bn:4d} {tiles:6d} {waves:6.2f} {fill:6.2f} {fl:7.2f}")
    print(f"\nGEMM work {tot_ideal:.1f} TFLOP per forward; FLOP-weighted tile fill {tot_ideal / tot_model:.3f} "
          f"-> a scheduler without wave quantisation / column padding would shorten the GEMM part by {(1 - tot_ideal / tot_model) * 100:.1f} %")


if __name__ == "__main__":
    main()
