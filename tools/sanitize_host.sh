#!/bin/bash
# Host-side loader / flatten / engine under ASan + UBSan and TSan (no GPU, no CUDA headers needed).
# usage: tools/sanitize_host.sh <edge list file>      (e.g. the file bench.graph_file("youtube") writes)
set -eu
ROOT=$(cd "$(dirname "$0")/.." && pwd); CS=$ROOT/graphvite_b200/csrc; OUT=${TMPDIR:-/tmp}/gv_sanitize; mkdir -p $OUT
cat > $OUT/graph_main.cpp <<'CPP'
#include "gv_host.h"
#include <cstdio>
int main(int argc, char **argv) {
    gv::Graph g;
    g.load_file(argv[1], true, false, " \t\r\n", "#");
    float p;
    const bool uniform = g.uniform_edge_table(p);
    std::printf("%u vertices, %llu lines, %zu directed edges, uniform %d\n", g.num_vertex(), (unsigned long long)g.num_edge, g.edge_u.size(), int(uniform));
    gv::Graph h;
    h.load_file(argv[1], false, true, " \t\r\n", "#");
    std::printf("directed + normalized: %zu edges\n", h.edge_u.size());
    return 0;
}
CPP
cat > $OUT/engine_main.cpp <<'CPP'
#include "gv_engine.h"
#include <cstdio>
#include <vector>
int main() {
    gv::Mt19937 parallel(7), sequential(7);
    std::vector<float> a(size_t(gv::kJumpBlocks) * 624 * 5 + 1234), b(a.size());
    parallel.fill_uniform(a.data(), a.size(), -1.f, 1.f, 4);
    sequential.fill_uniform(b.data(), b.size(), -1.f, 1.f, 1);
    std::printf("identical draws %d, identical state %d\n", int(a == b), int(parallel() == sequential()));
    return !(a == b);
}
CPP
for SAN in address,undefined thread; do
  g++ -std=c++17 -O1 -g -fsanitize=$SAN -fno-omit-frame-pointer -I$CS -I$ROOT/include $OUT/graph_main.cpp $CS/gv_graph.cpp $CS/gv_error.cpp -lpthread -o $OUT/graph_$SAN
  g++ -std=c++17 -O1 -g -fsanitize=$SAN -I$CS $OUT/engine_main.cpp -lpthread -o $OUT/engine_$SAN
  echo "== $SAN"; $OUT/graph_$SAN "$1"; GV_LOAD_CHUNK=37 $OUT/graph_$SAN "$1" | tail -1; $OUT/engine_$SAN
done
