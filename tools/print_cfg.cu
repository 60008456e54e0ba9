#include "rs_kernels.cuh"
#include <cstdio>
using namespace garage_ec;
template <int K, int MODE> void pr() {
    using C = StreamCfg<K, MODE>;
    printf("K=%2d mode=%d tma=%d S=%2d groups=%d [", K, MODE, (int)C::kTma, C::S, C::kLay.ngroups);
    for (int i = 0; i < C::kLay.ngroups; i++) printf("%d ", 1 << C::kLay.lg[i]);
    printf("] tab=%3uKB stage=%3uKB smem=%3uKB warps=%d(+%d) threads=%d regcap=%d\n", C::kTabBytes >> 10, C::kStageBytes >> 10, C::kSmem >> 10,
           C::kWarps, C::kWarpsAll - C::kWarps, C::kThreads, 65536 / C::kThreads);
}
template <int K> void prk() { pr<K,0>(); pr<K,1>(); pr<K,2>(); }
int main() { prk<4>(); prk<6>(); prk<7>(); prk<8>(); prk<10>(); prk<12>(); prk<13>(); prk<14>(); prk<15>(); prk<16>(); prk<17>(); prk<20>(); prk<24>(); prk<25>(); prk<28>(); prk<29>(); prk<32>(); }
