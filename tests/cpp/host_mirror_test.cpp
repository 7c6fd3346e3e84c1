// Compile + link check of include/poseidon252_b200.hpp against the C-ABI library, plus the host-only
// behaviour (domain separators, io-pattern errors, no-GPU failure).  Built and run by tests/test_cpp_host.py.
#include <cstdio>
#include <cstring>

#include "poseidon252_b200.hpp"

int main() {
    using namespace p252;
    if (domain_separator(Domain::Merkle4) != 0xf || domain_separator(Domain::Merkle2) != 0x3 ||
        domain_separator(Domain::Encryption) != 0x100000000ull || domain_separator(Domain::Other) != 0)
        return 1;
    Scalar tag;
    if (p252_hash_tag(P252_DOMAIN_MERKLE4, 3, 1, &tag) != P252_ERR_IO_PATTERN_VIOLATION) return 2;
    int ndev = 0;
    p252_device_count(&ndev);
    std::vector<Scalar> four(4);
    memset(four.data(), 0, sizeof(Scalar) * 4);
    try {
        Hash h(Domain::Merkle4);
        h.update(four.data(), 3);
        h.finalize();
        return 3;   // must not get here: Merkle4 with 3 inputs
    } catch (const Error& e) {
        if (!e.is_io_pattern_violation()) return 4;
    }
    if (ndev == 0) {
        try {
            Hash::digest(Domain::Merkle4, four);
            return 5;   // no CPU fallback allowed
        } catch (const Error& e) {
            if (e.code != P252_ERR_NO_DEVICE) return 6;
        }
        std::puts("host mirror ok (no GPU: batch path refuses to run)");
        return 0;
    }
    // with a GPU: digest == digest_batch of one; chunked update == one-shot (README.md:40-47)
    std::vector<Scalar> in(42);
    for (size_t i = 0; i < in.size(); ++i) in[i] = Scalar{{i + 1, 0, 0, 0}};
    auto one = Hash::digest(Domain::Other, in);
    Hash h(Domain::Other);
    h.update(in.data(), 3);
    h.update(in.data() + 3, 39);
    auto two = h.finalize();
    if (memcmp(one.data(), two.data(), sizeof(Scalar)) != 0) return 7;
    auto b = Hash::digest_batch(Domain::Other, in.data(), 1, 42);
    if (memcmp(one.data(), b.data(), sizeof(Scalar)) != 0) return 8;
    Scalar uv[2] = {in[0], in[1]};
    auto c = encrypt({in[2], in[3], in[4]}, uv, in[5]);
    auto m = decrypt(c, uv, in[5]);
    if (m.size() != 3 || memcmp(m.data(), &in[2], 3 * sizeof(Scalar)) != 0) return 9;
    c[0].l[0] ^= 1;
    try {
        decrypt(c, uv, in[5]);
        return 10;
    } catch (const Error& e) {
        if (!e.is_decryption_failed()) return 11;
    }
    // Merkle tree + openings (arity 4 and 2): every opening verifies, a wrong item or a tampered sibling does not
    for (int arity : {4, 2}) {
        std::vector<Scalar> leaves(64);
        for (size_t i = 0; i < leaves.size(); ++i) leaves[i] = Scalar{{100 + i, i, 0, 0}};
        auto nodes = merkle_build(arity, leaves);
        auto ops = merkle_open_batch(arity, leaves, nodes, {0, 7, 63});
        if (ops.size() != 3 || ops[1].depth() != (arity == 4 ? 3u : 6u)) return 12;
        if (!ops[1].verify(leaves[7]) || ops[1].verify(leaves[8])) return 13;
        Opening bad = ops[2];
        bad.branch[1].l[1] ^= 4;
        if (!ops[2].verify(leaves[63]) || bad.verify(leaves[63])) return 14;
    }
    std::puts("host mirror ok (GPU)");
    return 0;
}
