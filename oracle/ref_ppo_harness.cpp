// oracle/ref_ppo_harness.cpp — TEST INFRASTRUCTURE ONLY.
//
// C ABI over the UNMODIFIED reference's PPO-side arithmetic, compiled in place against the pip libtorch
// (oracle/Makefile target ref_ppo -> oracle/_ref/librlref_ppo.so):
//   RLGPC::TorchFuncs::ComputeGAE        RLGymPPO_CPP/src/private/RLGymPPO_CPP/Util/TorchFuncs.cpp:5-52
//   RLGPC::DiscretePolicy                RLGymPPO_CPP/src/private/RLGymPPO_CPP/PPO/DiscretePolicy.cpp:7-75
//   RLGPC::ValueEstimator                RLGymPPO_CPP/src/private/RLGymPPO_CPP/PPO/ValueEstimator.cpp:6-27
//   RLGPC::ExperienceBuffer              RLGymPPO_CPP/src/private/RLGymPPO_CPP/PPO/ExperienceBuffer.cpp:12-121
// It pins oracle/ppo_oracle.py (and, through it, the device kernels) to what the reference's own binaries compute.
#include <private/RLGymPPO_CPP/PPO/DiscretePolicy.h>
#include <private/RLGymPPO_CPP/PPO/ExperienceBuffer.h>
#include <private/RLGymPPO_CPP/PPO/ValueEstimator.h>
#include <private/RLGymPPO_CPP/Util/TorchFuncs.h>

#include <torch/torch.h>

#include <cstring>

using namespace RLGPC;

extern "C" {

// rews/dones/truncated: [n]; values: [n + 1]; outputs: [n]
int ref_ppo_gae(int n, const float* rews, const float* dones, const float* truncated, const float* values, float gamma, float lambda,
                float returnStd, float clipRange, float* outAdv, float* outValueTargets, float* outReturns) {
    try {
        FList r(rews, rews + n), d(dones, dones + n), t(truncated, truncated + n), v(values, values + n + 1), ret;
        torch::Tensor adv, tgt;
        TorchFuncs::ComputeGAE(r, d, t, v, adv, tgt, ret, gamma, lambda, returnStd, clipRange);
        adv = adv.cpu().contiguous().to(torch::kFloat32);
        tgt = tgt.cpu().contiguous().to(torch::kFloat32);
        if (adv.numel() != n || tgt.numel() != n || (int)ret.size() != n) return -2;
        memcpy(outAdv, adv.data_ptr<float>(), n * 4);
        memcpy(outValueTargets, tgt.data_ptr<float>(), n * 4);
        memcpy(outReturns, ret.data(), n * 4);
        return 0;
    } catch (std::exception& e) {
        fprintf(stderr, "ref_ppo_gae: %s\n", e.what());
        return -1;
    }
}

static torch::Tensor blob(const void* p, std::vector<int64_t> shape, torch::ScalarType t = torch::kFloat32) {
    return torch::from_blob(const_cast<void*>(p), shape, torch::TensorOptions().dtype(t));
}
// parameters() of Sequential(Linear, ReLU, ..., Linear) = weight0, bias0, weight1, bias1, ...
static void load_params(torch::nn::Sequential seq, const float* const* W, const float* const* b) {
    torch::NoGradGuard ng;
    auto params = seq->parameters();
    for (size_t i = 0; i < params.size(); i++) {
        auto& p = params[i];
        const float* src = (i & 1) ? b[i / 2] : W[i / 2];
        p.copy_(blob(src, p.sizes().vec()));
    }
}

// DiscretePolicy with the given nn.Linear tensors (W[l]: [out, in] row-major): GetActionProbs, the deterministic GetAction and
// GetBackpropData (log-prob of acts, mean entropy) on obs [rows, in]
int ref_ppo_policy(int rows, int in, int nActions, int nHidden, const int32_t* hidden, const float* const* W, const float* const* b,
                   float temperature, const float* obs, const int64_t* acts, float* outProbs, int64_t* outArgmax, float* outLogProbs,
                   float* outEntropy) {
    try {
        IList sizes(hidden, hidden + nHidden);
        DiscretePolicy policy(in, nActions, sizes, torch::kCPU, temperature);
        load_params(policy.seq, W, b);
        torch::NoGradGuard ng;
        auto o = blob(obs, {rows, in}).clone();
        auto probs = policy.GetActionProbs(o).contiguous();
        memcpy(outProbs, probs.data_ptr<float>(), (size_t)rows * nActions * 4);
        auto act = policy.GetAction(o, true).action.to(torch::kInt64).contiguous();
        memcpy(outArgmax, act.data_ptr<int64_t>(), (size_t)rows * 8);
        auto a = blob(acts, {rows, 1}, torch::kInt64).clone();
        auto bp = policy.GetBackpropData(o, a);
        auto lp = bp.actionLogProbs.cpu().contiguous().view({-1});
        memcpy(outLogProbs, lp.data_ptr<float>(), (size_t)rows * 4);
        *outEntropy = bp.entropy.item<float>();
        return 0;
    } catch (std::exception& e) {
        fprintf(stderr, "ref_ppo_policy: %s\n", e.what());
        return -1;
    }
}

int ref_ppo_critic(int rows, int in, int nHidden, const int32_t* hidden, const float* const* W, const float* const* b, const float* obs,
                   float* outValues) {
    try {
        IList sizes(hidden, hidden + nHidden);
        ValueEstimator critic(in, sizes, torch::kCPU);
        load_params(critic.seq, W, b);
        torch::NoGradGuard ng;
        auto o = blob(obs, {rows, in}).clone();
        auto v = critic.Forward(o).contiguous().view({-1});
        memcpy(outValues, v.data_ptr<float>(), (size_t)rows * 4);
        return 0;
    } catch (std::exception& e) {
        fprintf(stderr, "ref_ppo_critic: %s\n", e.what());
        return -1;
    }
}

// ExperienceBuffer(maxSize): SubmitExperience with `nSubmits` batches; batch k has sizes[k] rows whose `states` rows are
// [width] floats and whose other tensors carry one scalar per row, all filled with consecutive numbers from start[k] so the
// FIFO order is visible.  Returns curSize and the buffer's states / actions / advantages tensors ([maxSize, ...], NaN = unset).
int ref_ppo_buffer(int64_t maxSize, int width, int nSubmits, const int32_t* sizes, const float* start, float* outStates, float* outActions,
                   float* outAdvantages, int64_t* outCurSize, int64_t batchSize, int32_t* outNumBatches) {
    try {
        ExperienceBuffer buf(maxSize, 123, torch::kCPU);
        for (int k = 0; k < nSubmits; k++) {
            const int n = sizes[k];
            auto col = torch::arange(n, torch::kFloat32) + start[k];
            ExperienceTensors t;
            t.states = col.view({n, 1}).repeat({1, width}) + torch::arange(width, torch::kFloat32).view({1, width}) * 0.001f;
            t.actions = col + 0.1f; t.logProbs = col + 0.2f; t.rewards = col + 0.3f;
            t.nextStates = t.states + 0.5f; t.dones = col + 0.4f; t.truncated = col + 0.5f; t.values = col + 0.6f; t.advantages = col + 0.7f;
            buf.SubmitExperience(t);
        }
        *outCurSize = buf.curSize;
        memcpy(outStates, buf.data.states.contiguous().data_ptr<float>(), (size_t)maxSize * width * 4);
        memcpy(outActions, buf.data.actions.contiguous().data_ptr<float>(), (size_t)maxSize * 4);
        memcpy(outAdvantages, buf.data.advantages.contiguous().data_ptr<float>(), (size_t)maxSize * 4);
        *outNumBatches = batchSize > 0 ? (int32_t)buf.GetAllBatchesShuffled(batchSize).size() : 0;
        return 0;
    } catch (std::exception& e) {
        fprintf(stderr, "ref_ppo_buffer: %s\n", e.what());
        return -1;
    }
}
}
