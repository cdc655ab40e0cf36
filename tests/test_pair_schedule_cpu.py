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
    """Mirror of the epilogue's chunk loop (gemm_pair.cuh): yields (gk, t, c) for the chunks warp group `eg` processes."""
    gk = 0
    for t, _, nch in tiles:
        c = ((eg - gk) % G + G) % G
        while c < nch:
            yield gk + c, t, c
            c += G
        gk += nch


def test_groups_partition_the_chunk_sequence():
    rng = random.Random(3)
    for _ in range(300):
        G = rng.choice([2, 3, 4])
        tiles = [(7 + 13 * i, 0, rng.choice([1, 2, 3, 4, 5, 8])) for i in range(rng.randrange(1, 12))]
        total_chunks = sum(n for _, _, n in tiles)
        seen = {}
        for eg in range(G):
            last = -1
            for gk, t, c in _owned_chunks(eg, G, tiles):
                assert gk % G == eg and gk > last
                last = gk
                assert gk not in seen
                seen[gk] = (t, c)
        assert sorted(seen) == list(range(total_chunks))
        k = 0  # the manager numbers chunks tile after tile
        for t, _, nch in tiles:
            for c in range(nch):
                assert seen[k] == (t, c)
                k += 1


def test_constant_buffer_handshake_parities():
    """const_full / const_empty (two stages): the manager loads tile ordinal i into stage i & 1 after waiting for the
    ((i >> 1) - 1)-th completion of const_empty[i & 1]; the epilogue waits for completion (i >> 1) of const_full[i & 1]."""
    for n_tiles in range(1, 9):
        full = [0, 0]
        empty = [0, 0]
        loaded = []
        for i in range(min(2, n_tiles)):
            full[i & 1] += 1
            loaded.append(i)
        for it in range(n_tiles):          # epilogue consumes tile `it`, then the manager may load it + 2
            assert it in loaded and full[it & 1] == (it >> 1) + 1
            empty[it & 1] += 1
            nxt = it + 2
            if nxt < n_tiles:
                assert empty[nxt & 1] >= (nxt >> 1)  # the completion the manager waits for has happened
                full[nxt & 1] += 1
                loaded.append(nxt)
