// =============================================================================
// oracle/gv_oracle_common.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Pieces shared by the CPU restatements of the reference (oracle/gv_oracle.cpp: node embedding,
// oracle/gv_oracle_kg.cpp: knowledge-graph embedding): the process-wide engine, AliasTable,
// partition(), the optimizers and the cuRAND host stream.  See gv_oracle.cpp for the parity status.
// =============================================================================
#pragma once

#include <curand.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <queue>
#include <random>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace oracle {

typedef uint32_t Index;

// core/solver.h:51-57
static const int kMaxPartition = 16;
static const int kRandBatchSize = 5000000;
static const int kSamplePerVertex = 175;
static const int kMinEpisodeSample = 20000000;
// instance/graph.cuh:56
static const int kExpectedDegree = 1600;
// util/common.h:28
static const float kEpsilon = 1e-15f;

inline void fail(const std::string &msg) {
    throw std::runtime_error(msg);
}

// -----------------------------------------------------------------------------
// R1: process-wide engine, core/solver.h:50 (default-constructed mt19937)
// -----------------------------------------------------------------------------
inline std::mt19937 &global_engine() {
    static std::mt19937 seed;
    return seed;
}

// -----------------------------------------------------------------------------
// R2/R3: AliasTable, base/alias_table.cuh:84-152
// -----------------------------------------------------------------------------
template<class I>
struct AliasTable {
    std::vector<float> prob;
    std::vector<I> alias;
    I count = 0;

    // base/alias_table.cuh:84-128
    void build(const std::vector<float> &weights) {
        count = weights.size();
        if (count == 0)
            fail("Invalid sampling distribution");
        prob = weights;
        alias.assign(count, 0);
        double norm = 0;  // :92 "single precision may cause considerable truncation error"
        for (size_t i = 0; i < count; i++)
            norm += prob[i];
        norm = norm / count;
        for (size_t i = 0; i < count; i++)
            prob[i] = float(double(prob[i]) / norm);  // float /= double

        std::queue<I> large, little;
        for (size_t i = 0; i < count; i++) {
            if (prob[i] < 1)
                little.push(i);
            else
                large.push(i);
        }
        while (!little.empty() && !large.empty()) {
            I i = little.front(), j = large.front();
            little.pop();
            large.pop();
            alias[i] = j;
            float t = prob[i] + prob[j];
            prob[j] = t - 1;
            if (prob[j] < 1)
                little.push(j);
            else
                large.push(j);
        }
        while (!little.empty()) {
            I i = little.front();
            little.pop();
            alias[i] = i;
        }
        while (!large.empty()) {
            I i = large.front();
            large.pop();
            alias[i] = i;
        }
    }

    // base/alias_table.cuh:148-152.  cuRAND doubles lie in (0,1], so rand1*count can
    // equal count (an out-of-bounds read in the reference); we clamp to count-1 and
    // change no other outcome.
    I sample(double rand1, double rand2) const {
        I index = I(rand1 * count);
        if (index >= count)
            index = count - 1;
        float p = float(rand2);
        return p < prob[index] ? index : alias[index];
    }
};

// -----------------------------------------------------------------------------
// R6: partition, core/solver.h:873-887 (unstable std::sort: tie order is libstdc++'s)
// -----------------------------------------------------------------------------
inline std::vector<std::vector<Index>> partition(const std::vector<float> &weights, int num_partition) {
    std::vector<Index> indexes(weights.size());
    for (Index i = 0; i < indexes.size(); i++)
        indexes[i] = i;
    std::sort(indexes.begin(), indexes.end(), [&weights](Index x, Index y) { return weights[x] > weights[y]; });
    std::vector<std::vector<Index>> parts(num_partition);
    for (Index i = 0; i < indexes.size(); i++) {
        int part_id = i % (num_partition * 2);
        part_id = std::min(part_id, num_partition * 2 - 1 - part_id);
        parts[part_id].push_back(indexes[i]);
    }
    return parts;
}

// -----------------------------------------------------------------------------
// R20: optimizer, core/optimizer.h:42-85,132-134,161-210
// -----------------------------------------------------------------------------
enum OptimizerType { kSGD = 0, kMomentum, kAdaGrad, kRMSprop, kAdam };
enum ScheduleType { kConstant = 0, kLinear = 1 };

struct Optimizer {
    int type = kSGD;
    int schedule = kLinear;
    float init_lr = 0.025f, lr = 0.025f, weight_decay = 0.005f;
    float a = 0, b = 0;  // momentum | alpha | beta1, beta2
    float epsilon = 0;

    int num_moment() const {
        return type == kSGD ? 0 : (type == kAdam ? 2 : 1);
    }
    // optimizer.h:77-85,132-134
    void apply_schedule(int batch_id, int num_batch) {
        float factor = 1;
        if (schedule == kLinear)
            factor = std::max(1 - float(batch_id) / num_batch, 1e-4f);
        lr = init_lr * factor;
    }
    // optimizer.h:161-164
    float sgd_update(float parameter, float gradient, float weight) const {
        return lr * weight * (gradient + weight_decay * parameter);
    }
    // optimizer.h:171-175
    float momentum_update(float parameter, float gradient, float &moment1, float weight) const {
        float regularized = weight * (gradient + weight_decay * parameter);
        moment1 = a * moment1 + (1 - a) * regularized;
        return lr * moment1;
    }
    // optimizer.h:182-186
    float adagrad_update(float parameter, float gradient, float &moment1, float weight) const {
        float regularized = weight * (gradient + weight_decay * parameter);
        moment1 += regularized * regularized;
        return lr * regularized / (sqrtf(moment1) + epsilon);
    }
    // optimizer.h:193-197
    float rmsprop_update(float parameter, float gradient, float &moment1, float weight) const {
        float regularized = weight * (gradient + weight_decay * parameter);
        moment1 = a * moment1 + (1 - a) * regularized * regularized;
        return lr * regularized / sqrtf(moment1 + epsilon);
    }
    // optimizer.h:203-210 (no bias correction)
    float adam_update(float parameter, float gradient, float &moment1, float &moment2, float weight) const {
        float regularized = weight * (gradient + weight_decay * parameter);
        moment1 = a * moment1 + (1 - a) * regularized;
        moment2 = b * moment2 + (1 - b) * regularized * regularized;
        return lr * moment1 / (sqrtf(moment2) + epsilon);
    }
    float update(float parameter, float gradient, float *m1, float *m2, float weight) const {
        switch (type) {
            case kSGD: return sgd_update(parameter, gradient, weight);
            case kMomentum: return momentum_update(parameter, gradient, *m1, weight);
            case kAdaGrad: return adagrad_update(parameter, gradient, *m1, weight);
            case kRMSprop: return rmsprop_update(parameter, gradient, *m1, weight);
            default: return adam_update(parameter, gradient, *m1, *m2, weight);
        }
    }
};

// util/math.h:30-33
inline float sigmoid(float x) {
    return x > 0 ? 1 / (1 + expf(-x)) : expf(x) / (expf(x) + 1);
}

// -----------------------------------------------------------------------------
// cuRAND host generator wrapper (XORWOW, same seeding calls as core/solver.h:950-953)
// -----------------------------------------------------------------------------
struct RandomStream {
    curandGenerator_t generator = nullptr;
    RandomStream(unsigned long long seed) {
        if (curandCreateGeneratorHost(&generator, CURAND_RNG_PSEUDO_DEFAULT) != CURAND_STATUS_SUCCESS)
            fail("curandCreateGeneratorHost failed");
        curandSetPseudoRandomGeneratorSeed(generator, seed);
    }
    ~RandomStream() {
        if (generator)
            curandDestroyGenerator(generator);
    }
    RandomStream(const RandomStream &) = delete;
    void generate(double *out, size_t n) {
        if (curandGenerateUniformDouble(generator, out, n) != CURAND_STATUS_SUCCESS)
            fail("curandGenerateUniformDouble failed");
    }
};


}  // namespace oracle
