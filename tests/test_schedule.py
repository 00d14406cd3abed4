"""Host logic of the persistent sweep's work table (fv2d_capi.cu: schedule_runs, through the
fv2d_debug_schedule test hook; no GPU needed): the runs tile the slab's rows exactly once, no run is
shorter than 8 rows unless the slab is (the kernel stages rows of the NEXT item while it finishes
the current one and never looks further), runs of a round have equal heights, and a slab with
neighbour slabs starts with its two short edge runs."""
import pytest

from fv2d_b200 import capi

CASES = [(8192, 8192, False, False), (4096, 4096, False, False), (8192, 1024, True, True), (16384, 16384, False, False),
         (16384, 2048, True, False), (64, 16, False, False), (300, 5, False, False), (777, 336, True, True),
         (252, 1, False, False), (100000, 40, False, True), (8192, 4096, False, True), (33, 100001, False, False)]


@pytest.mark.parametrize("Nx,Ny,lo,hi", CASES)
def test_runs_tile_the_slab(Nx, Ny, lo, hi):
    runs = capi.schedule_runs(Nx, Ny, 148, lo, hi)
    rows = sorted(runs)
    assert rows[0][0] == 0 and rows[-1][1] == Ny
    for (a0, a1), (b0, b1) in zip(rows, rows[1:]):
        assert a1 == b0 and a1 > a0
    heights = [b - a for a, b in runs]
    if Ny >= 8:
        assert min(heights) >= 8, heights
    assert max(heights) <= max(96 + 7, 0) or Ny < 8
    if lo and Ny >= 32:
        assert runs[0] == (0, 8)
    if hi and Ny >= 32:
        assert (Ny - 8, Ny) in runs[:2]


def test_headline_schedule_has_few_items_per_cta_and_a_fine_tail():
    runs = capi.schedule_runs(8192, 8192, 148)
    nstrips = (8192 + 251) // 252
    items = len(runs) * nstrips
    assert 10 <= items / 296 <= 16  # ~1 % of per-item overhead at 96-row runs
    assert [b - a for a, b in runs[:8]] == [96] * 8
    assert max(b - a for a, b in runs[-10:]) <= 16  # the last CTAs finish within a few rows of each other


def test_runs_of_a_round_are_equal():
    runs = capi.schedule_runs(4096, 4096, 148)
    heights = [b - a for a, b in runs]
    per_round = (296 + 8) // 17
    for r in range(0, len(heights) - per_round, per_round):
        assert len(set(heights[r:r + per_round])) == 1, (r, heights[r:r + per_round])


@pytest.mark.parametrize("Ny,Ng,rows", [(8192, 2, 0), (8192, 2, 128), (64, 2, 16), (100, 3, 16), (5, 2, 16), (1000, 2, 256),
                                        (16384, 2, 512), (33, 2, 16), (31, 2, 16)])
def test_stream_blocks_cover_the_slab_and_respect_the_stencil(Ny, Ng, rows):
    """Row blocks of fv2d_advance_host_stream (fv2d_capi.cu: stream_blocks): uploads tile the domain rows,
    sweeps tile them too, and a block never sweeps a row whose upper neighbours (2 rows) are not
    resident yet - except the last block, whose upper neighbours are ghost rows."""
    blocks = capi.stream_blocks(Ny, Ng, rows)
    jbeg, jend = Ng, Ng + Ny
    assert blocks[0][0] == jbeg and blocks[0][2] == jbeg and blocks[-1][1] == jend and blocks[-1][3] == jend
    for (u0, u1, s0, s1), (v0, v1, t0, t1) in zip(blocks, blocks[1:]):
        assert u1 == v0 and s1 == t0
    for k, (u0, u1, s0, s1) in enumerate(blocks):
        assert u1 > u0 and s1 > s0
        if k < len(blocks) - 1:
            assert s1 + 2 <= u1          # rows s1, s1 + 1 (read by row s1 - 1) have arrived
        assert s0 - 2 >= jbeg - Ng        # rows below come from earlier blocks or are low ghost rows
    B = rows if rows > 0 else 256
    assert len(blocks) == max(1, Ny // max(16, B))
