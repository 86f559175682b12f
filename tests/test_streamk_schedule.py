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
