"""Host model of the index arithmetic inside gemm_pair_kernel (magicdrive_b200/csrc/gemm_pair.cuh): the staging-buffer
manager walks tiles with integer division, the epilogue warps with multiply-high reciprocals (div_magic, capi_gemm.cu).
Both must produce the same (tile -> N tile, chunk count, pixel row) sequence, otherwise the chunk numbering that pairs
`res_full` / `out_ready` barriers diverges and the kernel deadlocks (it did once for m_groups == 1: ceil(2^32 / 1) does not
fit 32 bits)."""
import random


def div_magic(d):
    return (((1 << 32) + d - 1) // d) & 0xFFFFFFFF


def fdiv(x, magic, d):
    return x if d == 1 else (x * magic) >> 32


def manager_tiles(cluster_id, n_clusters, m_groups, n_tiles, out_cols, out_per_tile, ch_tile):
    t = cluster_id
    while t < m_groups * n_tiles:
        nt = t // m_groups
        left = (out_cols - nt * out_per_tile) // 32
        yield t, nt, min(left, ch_tile)
        t += n_clusters


def epilogue_tiles(cluster_id, n_clusters, m_groups, n_tiles, out_cols, out_per_tile, ch_tile):
    mg_magic = div_magic(m_groups)
    t = cluster_id
    while t < m_groups * n_tiles:
        nt = fdiv(t, mg_magic, m_groups)
        left = (out_cols - nt * out_per_tile) // 32
        yield t, nt, min(left, ch_tile)
        t += n_clusters


def test_magic_division_matches_integer_division():
    rng = random.Random(0)
    for d in list(range(1, 260)) + [rng.randrange(260, 70000) for _ in range(200)]:
        m = div_magic(d)
        for x in list(range(0, 3000)) + [rng.randrange(0, (1 << 32) // d) for _ in range(500)]:
            if x * d < (1 << 32):
                assert fdiv(x, m, d) == x // d, (x, d)


def test_manager_and_epilogue_walk_the_same_chunks():
    rng = random.Random(1)
    for _ in range(400):
        ctas = rng.choice([1, 2])
        m_tiles = rng.choice([1, 2, 3, 5, 9, 33, 132, 4200])
        m_groups = (m_tiles + ctas - 1) // ctas
        block_n = rng.choice([64, 128, 160, 256])
        geglu = block_n == 256 and rng.random() < 0.5
        out_cols = 32 * rng.randrange(1, 90)
        n_out = out_cols * (2 if geglu else 1)
        n_tiles = (n_out + block_n - 1) // block_n
        out_per_tile = block_n // 2 if geglu else block_n
        ch_tile = block_n // 64 if geglu else block_n // 32
        n_clusters = min(m_groups * n_tiles, 148 // ctas)
        for cid in {0, n_clusters // 2, n_clusters - 1}:
            a = list(manager_tiles(cid, n_clusters, m_groups, n_tiles, out_cols, out_per_tile, ch_tile))
            b = list(epilogue_tiles(cid, n_clusters, m_groups, n_tiles, out_cols, out_per_tile, ch_tile))
            assert a == b
            assert all(nch >= 1 for _, _, nch in a)


def test_row_location_matches_reference_decomposition():
    rng = random.Random(2)
    for _ in range(300):
        ctas = rng.choice([1, 2])
        tiles_w, tiles_h, tiles_n = rng.choice([1, 1, 2, 4, 7]), rng.choice([1, 1, 2, 4, 7]), rng.choice([1, 2, 12, 96])
        m_tiles = tiles_w * tiles_h * tiles_n
        m_groups = (m_tiles + ctas - 1) // ctas
        n_tiles = rng.randrange(1, 9)
        mg, tw_m, th_m = div_magic(m_groups), div_magic(tiles_w), div_magic(tiles_h)
        for t in range(0, m_groups * n_tiles + 148, max(1, (m_groups * n_tiles) // 50)):
            for rank in range(ctas):
                nt = fdiv(t, mg, m_groups)
                mt = (t - nt * m_groups) * ctas + rank
                a = fdiv(mt, tw_m, tiles_w)
                tw = mt - a * tiles_w
                tn = fdiv(a, th_m, tiles_h)
                th = a - tn * tiles_h
                assert nt == t // m_groups
                assert (tw, th, tn) == (mt % tiles_w, (mt // tiles_w) % tiles_h, mt // (tiles_w * tiles_h))


def _owned_chunks(eg, G, tiles):
    """Mirror of the epilogue's chunk loop + one-chunk-ahead constant fetch (gemm_pair.cuh): yields (gk, t, c, prefetched)."""
    def first_owned(gk0, nch):
        c = ((eg - gk0) % G + G) % G
        return c if c < nch else -1

    gk = 0
    kc = None
    if tiles:
        c0 = first_owned(0, tiles[0][2])
        if c0 >= 0:
            kc = (tiles[0][0], c0)
    for i, (t, _, nch) in enumerate(tiles):
        gk_tile = gk
        c = first_owned(gk_tile, nch)
        while 0 <= c < nch:
            hit = kc == (t, c)
            yield gk_tile + c, t, c, hit
            if c + G < nch:
                kc = (t, c + G)
            elif i + 1 < len(tiles):
                cn = first_owned(gk_tile + nch, tiles[i + 1][2])
                if cn >= 0:
                    kc = (tiles[i + 1][0], cn)
            c += G
        gk = gk_tile + nch


def test_groups_partition_the_chunk_sequence_and_prefetch_hits():
    rng = random.Random(3)
    for _ in range(300):
        G = rng.choice([2, 3, 4])
        n_tiles_seq = rng.randrange(1, 12)
        tiles = [(7 + 13 * i, 0, rng.choice([1, 2, 3, 4, 5, 8])) for i in range(n_tiles_seq)]
        total_chunks = sum(n for _, _, n in tiles)
        seen = {}
        misses = 0
        for eg in range(G):
            last = -1
            for gk, t, c, hit in _owned_chunks(eg, G, tiles):
                assert gk % G == eg and gk > last
                last = gk
                assert gk not in seen
                seen[gk] = (t, c)
                misses += (not hit)
        assert sorted(seen) == list(range(total_chunks))
        # the manager numbers chunks tile after tile
        k = 0
        for t, _, nch in tiles:
            for c in range(nch):
                assert seen[k] == (t, c)
                k += 1
        # a fetch-at-use only happens after a tile in which the group owned nothing
        if all(n >= G for _, _, n in tiles):
            assert misses == 0
