"""Host-side model of the persistent GEMM's work-unit -> tile mapping (gemm_tcgen05.cuh: `tm = (n_fast ? unit / tiles_n : unit % tiles_mc) * CL + crank`,
`tn = n_fast ? unit % tiles_n : unit / tiles_mc`) and of the host's choice of the order (gemm.cu: N first when M > N and the tiles need more than one
wave): every (m tile, n tile) pair is produced exactly once by exactly one (unit, cluster rank) in both orders, and the order that is chosen is the one whose
modelled DRAM traffic (the swept operand is re-read once per wave unless it fits the L2) is the smaller one on the shapes of the engine."""
import math

import pytest

CL, BM, BN, CLUSTERS = 2, 128, 256, 74
L2_BYTES = 126 << 20


def tiles(M, N):
    tiles_m, tiles_n = math.ceil(M / BM), math.ceil(N / BN)
    return math.ceil(tiles_m / CL), tiles_n, tiles_m


def unit_to_tile(unit, crank, tiles_mc, tiles_n, n_fast):
    tm = (unit // tiles_n if n_fast else unit % tiles_mc) * CL + crank
    tn = unit % tiles_n if n_fast else unit // tiles_mc
    return tm, tn


def host_n_fast(M, N):
    tiles_mc, tiles_n, _ = tiles(M, N)
    return M > N and tiles_mc * tiles_n > CLUSTERS


def modelled_dram_bytes(M, N, K, n_fast):
    """operands only: the operand the units of a wave share is read once; the other one once per wave unless it stays in the L2"""
    tiles_mc, tiles_n, _ = tiles(M, N)
    waves = math.ceil(tiles_mc * tiles_n / CLUSTERS)
    a, b = M * K * 2, N * K * 2
    if n_fast:        # A row blocks read once; B swept by every wave
        return a + (b if b <= L2_BYTES // 2 else b * waves)
    return b + (a if a <= L2_BYTES // 2 else a * min(waves, tiles_n))


SHAPES = [(4128, 2048, 10240), (4128, 14336, 2048), (4096, 8192, 2048), (9240, 2048, 10240), (9240, 14336, 2048), (9240, 2048, 14336), (9240, 10240, 2048),
          (14336, 2048, 9240), (2048, 10240, 9240), (58498, 2048, 9240), (9240, 58498, 2048), (16416, 2048, 10240), (300, 520, 192), (257, 1000, 200)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_every_tile_once_in_both_orders(M, N, K):
    tiles_mc, tiles_n, tiles_m = tiles(M, N)
    for n_fast in (False, True):
        seen = set()
        for unit in range(tiles_mc * tiles_n):
            for crank in range(CL):
                tm, tn = unit_to_tile(unit, crank, tiles_mc, tiles_n, n_fast)
                assert 0 <= tn < tiles_n and 0 <= tm < tiles_mc * CL
                assert (tm, tn) not in seen
                seen.add((tm, tn))
        assert {(tm, tn) for tm in range(tiles_m) for tn in range(tiles_n)} <= seen       # (a padding m tile of the last pair is masked by the row bound)
    # consecutive units of the chosen order share the operand the order is named after
    nf = host_n_fast(M, N)
    if tiles_mc * tiles_n > 1:
        t0, t1 = unit_to_tile(0, 0, tiles_mc, tiles_n, nf), unit_to_tile(1, 0, tiles_mc, tiles_n, nf)
        if nf and tiles_n > 1:
            assert t0[0] == t1[0] and t0[1] != t1[1]          # same A row block, next B tile
        if not nf and tiles_mc > 1:
            assert t0[1] == t1[1] and t0[0] != t1[0]          # same B tile, next A row block


@pytest.mark.parametrize("M,N,K", [s for s in SHAPES if s[0] * s[1] > 256 * 256 * 74])
def test_chosen_order_is_the_cheaper_one_in_the_traffic_model(M, N, K):
    nf = host_n_fast(M, N)
    chosen, other = modelled_dram_bytes(M, N, K, nf), modelled_dram_bytes(M, N, K, not nf)
    assert chosen <= other, (M, N, K, nf, chosen, other)
