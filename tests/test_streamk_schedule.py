"""Host-side model of the stream-K schedule and hand-over protocol of the decode GEMM (show-o_b200/csrc/skinny.cuh:
sk2_begin / sk2_cta_of / sk2_consume) -- the arithmetic is mirrored line for line and its invariants are checked over many
(tiles, chunks per tile, grid) shapes, including the ones the engine launches.  (The grid > total case is the one the decode
megakernel first got wrong on the GPU: a CTA without work was counted as a contributor of a split tile.)"""
import itertools

import pytest


def sk2_begin(total, grid, cta):
    return total if cta >= grid else (cta * total) // grid


def sk2_cta_of(total, grid, chunk):
    return ((chunk + 1) * grid - 1) // total


def simulate(tiles, cpt, launch_grid):
    total = tiles * cpt
    grid = min(launch_grid, total)                # per-phase grid: every CTA below it owns >= 1 chunk
    covered = [0] * total
    tickets = [0] * tiles                         # partial segments that will take a ticket, per tile
    finished = [0] * tiles                        # direct (whole-tile) epilogues
    slots = {}
    for cta in range(launch_grid):
        b, e = sk2_begin(total, grid, cta), sk2_begin(total, grid, cta + 1)
        assert b <= e
        if cta < grid:
            assert e > b, "a CTA below the phase grid owns no chunk"
        else:
            assert e == b == total
        c, deferred = b, 0
        while c < e:
            tile = c // cpt
            t0 = tile * cpt
            seg_end = min(t0 + cpt, e)
            for k in range(c, seg_end):
                covered[k] += 1
            whole = c == t0 and seg_end == t0 + cpt
            if whole:
                finished[tile] += 1
            else:
                slot = (cta, 1 if c == t0 else 0)
                assert slot not in slots, "two partial segments of one CTA share a workspace slot"
                slots[slot] = tile
                tickets[tile] += 1
                if seg_end < e:
                    deferred += 1                 # ticket taken in the background, looked at after the last chunk
            c = seg_end
        assert deferred <= 1
    assert all(v == 1 for v in covered), "every chunk is contracted exactly once"
    for tile in range(tiles):
        t0 = tile * cpt
        first, last = sk2_cta_of(total, grid, t0), sk2_cta_of(total, grid, t0 + cpt - 1)
        assert sk2_begin(total, grid, first) <= t0 < sk2_begin(total, grid, first + 1)
        assert sk2_begin(total, grid, last) <= t0 + cpt - 1 < sk2_begin(total, grid, last + 1)
        if first == last:
            assert finished[tile] == 1 and tickets[tile] == 0
        else:
            # the last arriver is recognised by ticket == last - first: exactly that many + 1 CTAs must contribute
            assert finished[tile] == 0 and tickets[tile] == last - first + 1
            for cta in range(first, last + 1):    # the reader's slot choice matches the writer's
                b0 = sk2_begin(total, grid, cta)
                assert slots[(cta, 1 if b0 <= t0 else 0)] == tile


ENGINE_SHAPES = [(224, 16), (32, 80), (915, 16), (28, 2), (4, 10), (1, 1), (16, 1)]       # W1, W2, head, tiny model, ...


@pytest.mark.parametrize("tiles,cpt", ENGINE_SHAPES)
@pytest.mark.parametrize("grid", [1, 2, 7, 132, 148, 160])
def test_engine_shapes(tiles, cpt, grid):
    simulate(tiles, cpt, grid)


def test_sweep_small_shapes():
    for tiles, cpt, grid in itertools.product(range(1, 14), range(1, 12), (1, 2, 3, 5, 8, 13, 148)):
        simulate(tiles, cpt, grid)


# ------------------------------------------------------------------------------------------------ decode megakernel
def mega_items(cta, launch_grid, NL, W1N, D, F, M, H, n_keys):
    """The ring items of one CTA of decode_mega_kernel in CONSUMPTION order (show-o_b200/csrc/decode_mega.cu): per layer the
    GEMM1 chunks of its stream-K range, the 64-key chunks of its attention units, the GEMM2 chunks."""
    t1, c1 = W1N // 64, D // 128
    t2, c2 = D // 64, (D + F) // 128
    g1, g2 = min(launch_grid, t1 * c1), min(launch_grid, t2 * c2)
    b1, e1 = sk2_begin(t1 * c1, g1, cta), sk2_begin(t1 * c1, g1, cta + 1)
    b2, e2 = sk2_begin(t2 * c2, g2, cta), sk2_begin(t2 * c2, g2, cta + 1)
    units = list(range(cta, M * H, launch_grid))
    n_chunks = (n_keys + 63) // 64
    out = []
    for l in range(NL):
        out += [(l, 0, c // c1, c % c1) for c in range(b1, e1)]
        out += [(l, 1, u, j) for u in units for j in range(n_chunks)]
        out += [(l, 2, c // c2, c % c2) for c in range(b2, e2)]
    return out, (b1, e1, c1, b2, e2, c2, len(units), n_chunks)


def producer_sequence(cta, launch_grid, NL, geo):
    """The producer's cursor walk (enter / advance lambdas of the kernel), re-stated with the same state variables."""
    b1, e1, c1, b2, e2, c2, my_units, n_chunks = geo
    cnt = [e1 - b1, my_units * n_chunks, e2 - b2]
    if sum(cnt) == 0:
        return []
    st = dict(layer=0, ph=0, idx=0, tile=0, kc=0)

    def enter():
        while st["layer"] < NL and cnt[st["ph"]] == 0:
            st["ph"] += 1
            if st["ph"] == 3:
                st["ph"], st["layer"] = 0, st["layer"] + 1
        st["idx"] = 0
        if st["ph"] == 0:
            st["tile"], st["kc"] = b1 // c1, b1 % c1
        elif st["ph"] == 2:
            st["tile"], st["kc"] = b2 // c2, b2 % c2
        else:
            st["tile"], st["kc"] = 0, 0

    seq = []
    enter()
    while st["layer"] < NL:
        tile = cta + st["tile"] * launch_grid if st["ph"] == 1 else st["tile"]
        seq.append((st["layer"], st["ph"], tile, st["kc"]))
        lim = c1 if st["ph"] == 0 else (c2 if st["ph"] == 2 else n_chunks)
        st["kc"] += 1
        if st["kc"] == lim:
            st["kc"], st["tile"] = 0, st["tile"] + 1
        st["idx"] += 1
        if st["idx"] == cnt[st["ph"]]:
            st["ph"] += 1
            if st["ph"] == 3:
                st["ph"], st["layer"] = 0, st["layer"] + 1
            enter()
    return seq


@pytest.mark.parametrize("geom", [dict(NL=24, W1N=14336, D=2048, F=8192, M=16, H=32, n_keys=289),
                                  dict(NL=2, W1N=1792, D=256, F=1024, M=3, H=4, n_keys=271),
                                  dict(NL=3, W1N=1792, D=256, F=1024, M=1, H=4, n_keys=64)])
def test_megakernel_producer_issues_items_in_the_consumers_order(geom):
    grid = 148
    for cta in range(grid):
        items, geo = mega_items(cta, grid, **geom)
        assert producer_sequence(cta, grid, geom["NL"], geo) == items
    # the gated attention chunk (the one holding the current token) is the last chunk of every unit
    n_chunks = (geom["n_keys"] + 63) // 64
    assert (geom["n_keys"] - 1) // 64 == n_chunks - 1
