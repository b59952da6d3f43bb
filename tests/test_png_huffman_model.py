"""Line-by-line Python model of png_code_lengths (tray_b200/csrc/tray_png.cuh): two-queue Huffman tree over the symbols sorted
by (frequency, index), then zlib's gen_bitlen overflow redistribution. The device code must produce a COMPLETE prefix code
(Kraft sum exactly 1) of bounded length for any histogram, or inflate rejects the stream; the adversarial cases (Fibonacci
frequencies push the tree far beyond 15 levels) are checked here on the CPU, and on the GPU in tests/test_gpu_png.py."""
import random


def code_lengths(freq, order, m, maxbits, n_sym):
    len_out = [0] * n_sym
    node_w, parent, depth = [0] * (2 * m), [0] * (2 * m), [0] * (2 * m)
    for i in range(m):
        node_w[i] = freq[order[i]]
    li, ii, k = 0, m, m
    while k < 2 * m - 1:
        pick = [0, 0]
        for t in range(2):
            if li < m and (ii >= k or node_w[li] <= node_w[ii]):
                pick[t] = li
                li += 1
            else:
                pick[t] = ii
                ii += 1
        node_w[k] = node_w[pick[0]] + node_w[pick[1]]
        parent[pick[0]] = parent[pick[1]] = k
        k += 1
    bl, overflow = [0] * 32, 0
    for nd in range(2 * m - 3, -1, -1):
        d = depth[parent[nd]] + 1
        if d > maxbits:      # every node below the limit counts, interior ones too (zlib gen_bitlen)
            d = maxbits
            overflow += 1
        depth[nd] = d
        if nd < m:
            bl[d] += 1
    while overflow > 0:
        bits = maxbits - 1
        while bl[bits] == 0:
            bits -= 1
        bl[bits] -= 1
        bl[bits + 1] += 2
        bl[maxbits] -= 1
        overflow -= 2
    i = 0
    for bits in range(maxbits, 0, -1):
        for _ in range(bl[bits]):
            len_out[order[i]] = bits
            i += 1
    return len_out


def check(freq, maxbits):
    n = len(freq)
    used = [s for s in range(n) if freq[s] > 0]
    order = sorted(used, key=lambda s: (freq[s], s))
    L = code_lengths(freq, order, len(order), maxbits, n)
    assert all((l > 0) == (freq[s] > 0) for s, l in enumerate(L))
    assert max(L) <= maxbits
    assert sum(2 ** (maxbits - l) for l in L if l > 0) == 2 ** maxbits      # complete code
    for a in used:
        for b in used:
            if freq[a] > freq[b]:
                assert L[a] <= L[b]
    return L


FIB = [1, 1]
while len(FIB) < 257:
    FIB.append(FIB[-1] + FIB[-2])


def test_literal_code_is_complete_and_bounded_for_adversarial_histograms():
    check(FIB[:40] + [0] * 217, 15)          # tree depth 39 before limiting
    check(FIB[:257], 15)
    check([1] * 257, 15)
    check([1, 1] + [0] * 255, 15)
    check([10 ** 6] + [1] * 256, 15)
    rnd = random.Random(1)
    for _ in range(300):
        f = [0] * 257
        for s in rnd.sample(range(257), rnd.randint(2, 257)):
            f[s] = int(rnd.paretovariate(0.5)) if rnd.random() < 0.7 else rnd.randint(1, 5)
        f[256] = 1
        check(f, 15)


def test_code_length_code_is_complete_within_7_bits():
    check(FIB[:19], 7)
    rnd = random.Random(2)
    for _ in range(300):
        f = [0] * 19
        for s in rnd.sample(range(19), rnd.randint(2, 19)):
            f[s] = int(rnd.paretovariate(0.4))
        check(f, 7)
