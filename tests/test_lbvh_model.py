"""Python model of lbvh_tree_kernel (tray_b200/csrc/tray_lbvh.cuh, Karras 2012): per internal node its key range, split and
children, from sorted unique 64-bit keys. Structural validity is checked on many adversarial key sets: every internal node's
children partition its range, every key ends in exactly one single-key leaf, node 0 is the root, and the collapse rule
(ranges of <= 4 keys become one leaf) covers every key exactly once. The GPU tests check the same builder through images."""
import random


def clz64(x):
    return 64 - x.bit_length()


def delta(keys, m, a, b):
    if b < 0 or b >= m:
        return -1
    return clz64(keys[a] ^ keys[b])


def build(keys):
    m = len(keys)
    rng, child, parent = {}, {}, {}
    for i in range(m - 1):
        d = 1 if delta(keys, m, i, i + 1) - delta(keys, m, i, i - 1) >= 0 else -1
        dmin = delta(keys, m, i, i - d)
        lmax = 2
        while delta(keys, m, i, i + lmax * d) > dmin:
            lmax *= 2
        l, t = 0, lmax // 2
        while t >= 1:
            if delta(keys, m, i, i + (l + t) * d) > dmin:
                l += t
            t //= 2
        j = i + l * d
        dnode = delta(keys, m, i, j)
        s, t = 0, l
        while True:
            t = (t + 1) // 2
            if delta(keys, m, i, i + (s + t) * d) > dnode:
                s += t
            if t <= 1:
                break
        gamma = i + s * d + min(d, 0)
        first, last = min(i, j), max(i, j)
        left = (m - 1) + gamma if first == gamma else gamma
        right = (m - 1) + gamma + 1 if last == gamma + 1 else gamma + 1
        rng[i] = (first, last)
        child[i] = (left, right)
        for c in (left, right):
            assert c not in parent, "node with two parents"
            parent[c] = i
    return rng, child, parent


def node_range(n, m, rng):
    return rng[n] if n < m - 1 else (n - (m - 1), n - (m - 1))


def check(keys):
    keys = sorted(set(keys))
    m = len(keys)
    if m < 2:
        return
    rng, child, parent = build(keys)
    assert 0 not in parent and len(parent) == 2 * m - 2          # everything but the root has exactly one parent
    assert rng[0] == (0, m - 1)
    for i in range(m - 1):
        (l, r), (a, b) = child[i], rng[i]
        la, lb = node_range(l, m, rng)
        ra, rb = node_range(r, m, rng)
        assert la == a and lb + 1 == ra and rb == b, (i, (a, b), (la, lb), (ra, rb))   # children partition the range
    # collapse rule of lbvh_boxes_kernel: walking down from the root, a node whose range holds <= 4 keys is a leaf
    covered = []
    stack = [0]
    depth_max = 0
    depth = {0: 1}
    while stack:
        n = stack.pop()
        a, b = node_range(n, m, rng)
        if n >= m - 1 or b - a + 1 <= 4:
            covered += list(range(a, b + 1))
            continue
        for c in child[n]:
            depth[c] = depth[n] + 1
            depth_max = max(depth_max, depth[c])
            stack.append(c)
    assert sorted(covered) == list(range(m))
    return depth_max


def test_radix_tree_is_valid_for_adversarial_key_sets():
    rnd = random.Random(3)
    check([0, 1])
    check([5, 6, 7])
    check(list(range(1000)))                                    # consecutive integers
    check([1 << k for k in range(64)])                          # one key per bit length: a degenerate chain
    check([(1 << 63) | k for k in range(300)] + list(range(300)))
    check([(0xABCDEF << 40) | (k << 16) | k for k in range(2000)])   # identical Morton prefix, ids differ (one Morton cell)
    for _ in range(300):
        n = rnd.randint(2, 400)
        kind = rnd.random()
        if kind < 0.3:
            ks = [rnd.getrandbits(64) for _ in range(n)]
        elif kind < 0.6:   # clustered: few distinct 48-bit Morton codes, 16-bit ids below
            cells = [rnd.getrandbits(48) for _ in range(rnd.randint(1, 5))]
            ks = [(rnd.choice(cells) << 16) | i for i in range(n)]
        else:              # keys that differ only in low bits of the Morton part
            base = rnd.getrandbits(48) & ~0xFFF
            ks = [((base | rnd.getrandbits(12)) << 16) | i for i in range(n)]
        check(ks)


def test_depth_of_the_chain_case_exceeds_the_traversal_stack_and_is_detected():
    """One key per bit length gives a 64-level chain: the device build reports its depth and tray_scene_upload falls back to
    the host's median split when it is above 40 (the traversal stack holds 48 entries)."""
    assert check([1 << k for k in range(64)] + [3 << k for k in range(1, 60, 2)]) > 40
