// A std::mt19937-compatible engine (same state, same output sequence, usable with the std:: distributions) with a
// bulk path for the one place that draws hundreds of millions of numbers: init_embeddings
// (std::uniform_real_distribution<float>(a, b)(engine) for every element, instance/graph.cuh:724-731).
// fill_uniform() produces bit for bit what that loop produces with libstdc++ (gv_engine_self_check() compares the two;
// tests/test_host_runtime.py).  The stream is sequential by definition, but the generator is linear over GF(2):
// jump(l) skips kJumpBlocks * 2^l * 624 draws in ~1 ms (gv_mt_jump.h, tools/mt_jump_poly.py), so fill_uniform() cuts a
// long request into one contiguous piece per thread, and every thread advances its copy of the engine to its piece -- 1.46e8 draws (Youtube,
// d = 128) are 0.17 s on one core and sit on the critical path of GraphSolver.train().
#pragma once

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <emmintrin.h>  // SSE2, baseline x86-64
#include <thread>
#include <vector>

#include "gv_mt_jump.h"

namespace gv {

class Mt19937 {
public:
    typedef uint_fast32_t result_type;  // like std::mt19937: the distributions see the same type, min() and max()
    static constexpr result_type min() { return 0; }
    static constexpr result_type max() { return 0xFFFFFFFFu; }
    static constexpr uint32_t default_seed = 5489u;

    explicit Mt19937(uint32_t value = default_seed) { seed(value); }

    void seed(uint32_t value) {
        state[0] = value;
        for (int i = 1; i < kN; i++)
            state[i] = 1812433253u * (state[i - 1] ^ (state[i - 1] >> 30)) + uint32_t(i);
        index = kN;
    }

    result_type operator()() {
        if (index >= kN)
            twist();
        return temper(state[index++]);
    }

    // out[i] = std::uniform_real_distribution<float>(a, b)(*this), i = 0 .. n-1
    // `threads` = 0: as many as the machine has (at most 16); requests shorter than two spans stay on this thread
    void fill_uniform(float *out, size_t n, float a, float b, int threads = 0) {
        const size_t span = size_t(kJumpBlocks) * kN;
        if (threads == 0)
            threads = int(std::min<unsigned>(16, std::max<unsigned>(1, std::thread::hardware_concurrency())));
        const size_t spans = (n + span - 1) / span;
        if (threads <= 1 || spans < 2) {
            fill_uniform_sequential(out, n, a, b);
            return;
        }
        // thread t generates the draws of spans [first(t), first(t + 1)) from a copy of the engine that it advances
        // to its first span itself (one jump per set bit of first(t)): nothing is sequential but the draws of a thread
        const size_t workers = std::min<size_t>(size_t(threads), spans);
        std::vector<Mt19937> engines(workers, *this);
        auto first_span = [&](size_t t) { return t * spans / workers; };
        auto work = [&](size_t t) {
            engines[t].skip_spans(first_span(t));
            const size_t begin = first_span(t) * span, end = std::min(n, first_span(t + 1) * span);
            engines[t].fill_uniform_sequential(out + begin, end - begin, a, b);
        };
        std::vector<std::thread> pool;
        for (size_t t = 1; t < workers; t++)
            pool.emplace_back(work, t);
        work(0);
        for (auto &thread : pool)
            thread.join();
        *this = engines[workers - 1];
    }

    // this engine after count * kJumpBlocks * 624 more draws
    void skip_spans(size_t count) {
        for (int level = 0; level < kJumpLevels - 1; level++)
            if ((count >> level) & 1)
                jump(level);
        for (size_t rest = count >> (kJumpLevels - 1); rest > 0; rest--)
            jump(kJumpLevels - 1);
    }

    // this engine after kJumpBlocks * 2^level * 624 more draws (the position inside the current block is kept)
    void jump(int level) {
        // Horner over the window (x_k .. x_{k+623}) sliding through one long buffer: h <- T h is one step of the
        // recurrence appended behind the window, h <- h + s adds the start block to the window
        const uint32_t *polynomial = kJumpPolynomial[level];
        int degree = 19936;
        while (degree > 0 && !((polynomial[degree >> 5] >> (degree & 31)) & 1u))
            degree--;
        std::vector<uint32_t> buffer(size_t(kN) + size_t(degree) + 8, 0u);
        uint32_t *window = buffer.data();
        for (int i = degree; i >= 0; i--) {
            if (i != degree) {  // h <- T h
                window[kN] = twist1(window[0], window[1], window[kM]);
                window++;
            }
            if ((polynomial[i >> 5] >> (i & 31)) & 1u) {
                int j = 0;
                for (; j + 4 <= kN; j += 4)
                    _mm_storeu_si128(reinterpret_cast<__m128i *>(window + j),
                                     _mm_xor_si128(_mm_loadu_si128(reinterpret_cast<const __m128i *>(window + j)),
                                                   _mm_loadu_si128(reinterpret_cast<const __m128i *>(state + j))));
                for (; j < kN; j++)
                    window[j] ^= state[j];
            }
        }
        // the window now starts one block before the target; the 31 low bits of its first word are not part of the
        // 19937-bit state and are undefined: one ordinary block update produces the target block from valid bits only
        const int position = index;
        std::memcpy(state, window, kN * sizeof(uint32_t));
        twist();
        index = position;
    }

    void fill_uniform_sequential(float *out, size_t n, float a, float b) {
        const float range = b - a;
        size_t done = 0;
        while (done < n) {
            if (index >= kN)
                twist();
            size_t chunk = size_t(kN - index);
            if (chunk > n - done)
                chunk = n - done;
            convert(state + index, out + done, chunk, a, range);
            index += int(chunk);
            done += chunk;
        }
    }

private:
    static const int kN = 624, kM = 397;
    uint32_t state[kN + 4];  // + padding for unaligned 4-wide loads
    int index;

    static uint32_t temper(uint32_t y) {
        y ^= y >> 11;
        y ^= (y << 7) & 0x9D2C5680u;
        y ^= (y << 15) & 0xEFC60000u;
        y ^= y >> 18;
        return y;
    }

    static inline __m128i twist4(__m128i current, __m128i next, __m128i far) {
        const __m128i upper = _mm_set1_epi32(int(0x80000000u)), lower = _mm_set1_epi32(0x7FFFFFFF);
        const __m128i matrix = _mm_set1_epi32(int(0x9908B0DFu)), one = _mm_set1_epi32(1);
        const __m128i y = _mm_or_si128(_mm_and_si128(current, upper), _mm_and_si128(next, lower));
        const __m128i odd = _mm_sub_epi32(_mm_setzero_si128(), _mm_and_si128(y, one));  // all ones where y is odd
        return _mm_xor_si128(_mm_xor_si128(far, _mm_srli_epi32(y, 1)), _mm_and_si128(odd, matrix));
    }
    static inline uint32_t twist1(uint32_t current, uint32_t next, uint32_t far) {
        const uint32_t y = (current & 0x80000000u) | (next & 0x7FFFFFFFu);
        return far ^ (y >> 1) ^ ((y & 1) ? 0x9908B0DFu : 0u);
    }

    void twist() {
        static_assert(kN == 624 && kM == 397, "the loop bounds below are written out for MT19937");
        // state[i] <- f(old state[i], old state[i + 1], state[i + 397]): new values are never read back here
        for (int i = 0; i < 224; i += 4)
            _mm_storeu_si128(reinterpret_cast<__m128i *>(state + i),
                             twist4(_mm_loadu_si128(reinterpret_cast<const __m128i *>(state + i)),
                                    _mm_loadu_si128(reinterpret_cast<const __m128i *>(state + i + 1)),
                                    _mm_loadu_si128(reinterpret_cast<const __m128i *>(state + i + 397))));
        for (int i = 224; i < 227; i++)
            state[i] = twist1(state[i], state[i + 1], state[i + 397]);
        // state[i] <- f(old state[i], old state[i + 1], NEW state[i - 227]) (written at least 227 steps earlier);
        // 623 - 227 = 99 * 4, so the vector loop ends exactly before the last element
        for (int i = 227; i < 623; i += 4)
            _mm_storeu_si128(reinterpret_cast<__m128i *>(state + i),
                             twist4(_mm_loadu_si128(reinterpret_cast<const __m128i *>(state + i)),
                                    _mm_loadu_si128(reinterpret_cast<const __m128i *>(state + i + 1)),
                                    _mm_loadu_si128(reinterpret_cast<const __m128i *>(state + i - 227))));
        state[623] = twist1(state[623], state[0], state[396]);
        index = 0;
    }

    // generate_canonical<float, 24> of libstdc++: float(draw) / 2^32 (round to nearest even), clamped below 1,
    // then uniform_real_distribution's r * (b - a) + a as two separately rounded operations
    static void convert(const uint32_t *raw, float *out, size_t n, float a, float range) {
        size_t i = 0;
        const __m128i mask7 = _mm_set1_epi32(int(0x9D2C5680u)), mask15 = _mm_set1_epi32(int(0xEFC60000u));
        const __m128i low16 = _mm_set1_epi32(0xFFFF);
        const __m128 scale16 = _mm_set1_ps(65536.0f), inverse = _mm_set1_ps(2.3283064365386963e-10f);  // 2^-32
        const __m128 one = _mm_set1_ps(1.0f), below_one = _mm_set1_ps(0.99999994f);
        const __m128 va = _mm_set1_ps(a), vrange = _mm_set1_ps(range);
        for (; i + 4 <= n; i += 4) {
            __m128i y = _mm_loadu_si128(reinterpret_cast<const __m128i *>(raw + i));
            y = _mm_xor_si128(y, _mm_srli_epi32(y, 11));
            y = _mm_xor_si128(y, _mm_and_si128(_mm_slli_epi32(y, 7), mask7));
            y = _mm_xor_si128(y, _mm_and_si128(_mm_slli_epi32(y, 15), mask15));
            y = _mm_xor_si128(y, _mm_srli_epi32(y, 18));
            // unsigned 32-bit -> float with one rounding: high half * 65536 is exact, the sum rounds once
            const __m128 high = _mm_mul_ps(_mm_cvtepi32_ps(_mm_srli_epi32(y, 16)), scale16);
            const __m128 low = _mm_cvtepi32_ps(_mm_and_si128(y, low16));
            __m128 r = _mm_mul_ps(_mm_add_ps(high, low), inverse);
            const __m128 clamp = _mm_cmpge_ps(r, one);
            r = _mm_or_ps(_mm_and_ps(clamp, below_one), _mm_andnot_ps(clamp, r));
            _mm_storeu_ps(out + i, _mm_add_ps(_mm_mul_ps(r, vrange), va));
        }
        for (; i < n; i++) {
            float r = float(temper(raw[i])) * 2.3283064365386963e-10f;
            if (r >= 1.0f)
                r = 0.99999994f;
            volatile float product = r * range;  // no contraction into an fma
            out[i] = product + a;
        }
    }
};

}  // namespace gv
