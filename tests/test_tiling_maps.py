"""CPU: the index algebra of two kernels restated in numpy and checked for exact coverage.

1. CTA pairs of gconv_kernel (csrc/conv_fprop_dgrad.cu, CG = 2): pair tile -> (image, M tile of each rank, N tile),
   contiguous pair-tile ranges per cluster, and the InstanceNorm partial-sum slot of every (image, CTA, lane quarter):
   every tile is produced exactly once, the two tiles of a pair lie in the same image, no two CTAs share a slot and
   every slot index is below the count the sizing call returns (b200unet_conv_fprop_partials, no GPU needed).
2. The 32 x 32 x 9 tiles of sgd_flat_kernel (csrc/train_aux.cu): tile j of a [Cout, Cin, 3, 3] tensor reads 32 runs of
   288 contiguous OIHW elements and writes the fprop pack [Cout][3][3][Cin], the dgrad pack [Cin][3][3][Cout] and the
   parity-stacked stride-2 pack [4 Cin][4][Cout]: all elements covered once, packs equal to their definitions."""
import numpy as np
import pytest


def ceil_div(a, b):
    return -(-a // b)


def pair_grid(n, oh, ow, cout, bn, mt, units=74):
    tiles_w, tiles_h = ceil_div(ow, 8), ceil_div(oh, 16 * mt)
    per_img = tiles_w * tiles_h // 2 * (cout // bn)  # pair tiles per image
    total = per_img * n
    tiles_per = max(1, ceil_div(total, units))
    clusters = ceil_div(total, tiles_per)
    slots = 8 * (ceil_div(per_img, tiles_per) + 1)
    return tiles_w, tiles_h, tiles_per, clusters, slots


@pytest.mark.parametrize("n,oh,ow,cout,bn,mt", [
    (32, 64, 64, 256, 256, 1), (32, 32, 32, 512, 256, 1), (32, 128, 128, 128, 128, 2), (32, 256, 256, 64, 64, 4),
    (32, 128, 128, 384, 192, 1), (3, 40, 20, 128, 128, 2), (2, 32, 32, 256, 256, 1), (5, 70, 12, 64, 64, 4),
])
def test_cta_pair_tiles_cover_every_tile_once_and_slots_are_private(n, oh, ow, cout, bn, mt):
    from unet_implementations_b200 import _lib
    tiles_w, tiles_h, tiles_per, clusters, slots = pair_grid(n, oh, ow, cout, bn, mt)
    tiles_per_img, n_tiles = tiles_w * tiles_h, cout // bn
    if tiles_per_img % 2:
        pytest.skip("odd tile count: the host keeps single CTAs")
    total = tiles_per_img // 2 * n * n_tiles
    seen = np.zeros((n * tiles_per_img, n_tiles), dtype=np.int32)
    slot_owner = {}
    for cluster in range(clusters):
        lo, hi = min(cluster * tiles_per, total), min(cluster * tiles_per + tiles_per, total)
        for rank in range(2):
            imgs = set()
            for tile in range(lo, hi):                      # the kernel's loop over its contiguous pair-tile range
                m_tile = (tile // n_tiles) * 2 + rank       # m_tile_of()
                n_tile = tile % n_tiles
                img = m_tile // tiles_per_img
                assert img == ((tile // n_tiles) * 2) // tiles_per_img  # both ranks of a pair: the same image
                seen[m_tile, n_tile] += 1
                imgs.add(img)
            for img in imgs:                                # flush(): one slot per (image, CTA, lane quarter)
                b0 = (img * (tiles_per_img // 2) * n_tiles) // tiles_per
                for q in range(4):
                    slot = ((cluster - b0) * 2 + rank) * 4 + q
                    assert 0 <= slot < slots
                    assert slot_owner.setdefault((img, slot), (cluster, rank, q)) == (cluster, rank, q)
    assert (seen == 1).all()
    assert _lib.call("b200unet_conv_fprop_partials", n, oh, ow, cout) >= slots


@pytest.mark.parametrize("cout,cin,stride2", [(32, 32, False), (64, 32, True), (128, 64, True), (64, 192, False)])
def test_flat_optimizer_tiles_emit_the_packs(cout, cin, stride2):
    rng = np.random.default_rng(0)
    w = rng.standard_normal((cout, cin, 3, 3)).astype(np.float32)
    flat = w.reshape(-1)
    per_o = cin * 9
    wf = np.full((cout, 9, cin), np.nan, np.float32)
    wd = np.full((cin, 9, cout), np.nan, np.float32)
    ws = np.zeros((4 * cin, 4, cout), np.float32)
    ws_hit = np.zeros_like(ws, dtype=np.int32)
    touched = np.zeros(flat.size, np.int32)
    ct_n = cin // 32
    tiles = (cout // 32) * ct_n
    assert flat.size == tiles * 9216 and ceil_div(flat.size, 1024) == 9 * tiles  # the tensor owns 9 blocks per tile
    for j in range(tiles):
        o0, ci0 = (j // ct_n) * 32, (j % ct_n) * 32
        base = o0 * per_o + ci0 * 9
        st = np.zeros((32, 288), np.float32)
        for e in range(32 * 288):                            # read: 32 runs of 288 contiguous elements
            ol, rl = divmod(e, 288)
            i = base + ol * per_o + rl
            touched[i] += 1
            st[ol, rl] = flat[i]
        for e in range(32 * 288):                            # fprop pack, ci fastest
            cl, tap, ol = e & 31, (e >> 5) % 9, e // 288
            wf[o0 + ol, tap, ci0 + cl] = st[ol, cl * 9 + tap]
        for e in range(32 * 288):                            # dgrad packs, o fastest
            ol, tap, cl = e & 31, (e >> 5) % 9, e // 288
            v = st[ol, cl * 9 + tap]
            ci, o = ci0 + cl, o0 + ol
            wd[ci, tap, o] = v
            kh, kw = divmod(tap, 3)
            ph, dh = (0 if kh == 1 else 1), (1 if kh == 0 else 0)
            pw, dw = (0 if kw == 1 else 1), (1 if kw == 0 else 0)
            ws[(ph * 2 + pw) * cin + ci, dh * 2 + dw, o] = v
            ws_hit[(ph * 2 + pw) * cin + ci, dh * 2 + dw, o] += 1
    assert (touched == 1).all()
    assert np.array_equal(wf, w.transpose(0, 2, 3, 1).reshape(cout, 9, cin))
    assert np.array_equal(wd, w.transpose(1, 2, 3, 0).reshape(cin, 9, cout))
    if stride2:
        # definition of the parity-stacked pack (pack_s2_dgrad_weights_kernel): class (ph, pw), shift (dh, dw) holds
        # W[kh(ph, dh)][kw(pw, dw)] where the class uses that shift, zero elsewhere
        ref = np.zeros_like(ws)
        for ph in range(2):
            for pw in range(2):
                for dh in range(2):
                    for dw in range(2):
                        kh = (1 if dh == 0 else -1) if ph == 0 else (2 if dh == 0 else 0)
                        kw = (1 if dw == 0 else -1) if pw == 0 else (2 if dw == 0 else 0)
                        if kh >= 0 and kw >= 0:
                            ref[(ph * 2 + pw) * cin:(ph * 2 + pw + 1) * cin, dh * 2 + dw, :] = w[:, :, kh, kw].T
        assert np.array_equal(ws, ref)
        assert ws_hit.max() == 1  # every tap lands in exactly one (class, shift) block: the zero blocks are never written
