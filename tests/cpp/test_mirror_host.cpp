// host-only exercise of the C++ mirror's verifier / key-file half (no GPU): read vk + cvk files written by the Python
// oracle, verify a proof read from a file
#include <cstdio>
#include <fstream>
#include "zkb200.hpp"
int main(int argc, char **argv) {
    if (argc < 5) return 2;
    zkb::VerifierKeyData vk = zkb::read_vk_file(argv[1]);
    auto g2 = zkb::read_cvk_file(argv[2]);
    std::array<uint8_t, 802> proof{};
    std::ifstream f(argv[3], std::ios::binary);
    f.read(reinterpret_cast<char *>(proof.data()), 802);
    std::vector<zkb::Fr> pub(vk.pi_roots.size());
    std::ifstream g(argv[4], std::ios::binary);
    g.read(reinterpret_cast<char *>(pub.data()), pub.size() * 32);
    int rc = zkb::verify(vk, pub, proof, g2.first, g2.second);
    proof[500] ^= 1;
    int rc_bad = zkb::verify(vk, pub, proof, g2.first, g2.second);
    std::printf("n=%zu roots=%zu verify=%d tampered=%d\n", vk.n, vk.pi_roots.size(), rc, rc_bad);
    return rc == 0 && rc_bad != 0 ? 0 : 1;
}
