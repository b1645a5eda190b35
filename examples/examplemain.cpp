// examplemain.cpp — the reference's example app (T/examplemain.cpp:58-151) against the B200 engine: the SAME
// EnvCreateFunc (plugin objects, weights, tick skip) and LearnerConfig values, driven through the shim's
// ThreadAgentManager surface.  Collection only: the PPO update lives in rlgymppo_cpp_b200/learner.py.
//
//   g++ -std=c++17 -O2 -Iinclude examples/examplemain.cpp -o examplemain
//       -Lrlgymppo_cpp_b200/csrc -lrlgym_b200 -Wl,-rpath,$PWD/rlgymppo_cpp_b200/csrc   (one line)
//   python -m rlgymppo_cpp_b200.meshes --out collision_meshes && ./examplemain collision_meshes [iterations] [--custom-setter] [--low-gravity]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>

#include <rlgym_b200_shim.hpp>

using namespace RLGPC;  // RLGymPPO
using namespace RLGSC;  // RLGymSim

static bool g_customSetter = false;
static bool g_lowGravity = false;  // a MutatorConfig through Gym's last constructor argument (G/Gym.h:18)

// A user-defined StateSetter (runs on the host through Arena/Car/Ball proxies): ball dropped above midfield, cars on
// their own half facing the ball.
class MidfieldDropState : public StateSetter {
public:
    std::mt19937 rng{7};
    GameState ResetState(Arena* arena) override {
        std::uniform_real_distribution<float> ux(-2000.f, 2000.f), uz(300.f, 1200.f);
        BallState bs;
        bs.pos = Vec(ux(rng), 0.f, uz(rng));
        arena->ball->SetState(bs);
        for (Car* car : arena->GetCars()) {
            CarState cs;
            float side = car->team == Team::BLUE ? -1.f : 1.f;
            cs.pos = Vec(ux(rng), side * 2500.f, 17.f);
            cs.rotMat = Angle(side < 0 ? 1.5707963f : -1.5707963f, 0, 0).ToRotMat();
            cs.boost = 50.f;
            car->SetState(cs);
        }
        return GameState();
    }
};

// Create the RLGymSim environment for each of our games (identical to the reference's EnvCreateFunc)
EnvCreateResult EnvCreateFunc() {
    constexpr int TICK_SKIP = 8;
    constexpr float NO_TOUCH_TIMEOUT_SECS = 10.f;

    EventReward::WeightScales ev;
    ev.teamGoal = 1.f;
    ev.concede = -1.f;
    auto rewards = new CombinedReward({
        {new FaceBallReward(), 0.1f},
        {new VelocityPlayerToBallReward(), 0.5f},
        {new VelocityBallToGoalReward(), 1.0f},
        {new EventReward(ev), 50.f},
    }, true);

    std::vector<TerminalCondition*> terminalConditions = {new NoTouchCondition(NO_TOUCH_TIMEOUT_SECS * 120 / TICK_SKIP), new GoalScoreCondition()};

    auto obs = new DefaultOBS();
    auto actionParser = new DiscreteAction();
    StateSetter* stateSetter = g_customSetter ? (StateSetter*)new MidfieldDropState() : (StateSetter*)new RandomState(true, true, true);

    Match* match = new Match(rewards, terminalConditions, obs, actionParser, stateSetter, 1, true);
    MutatorConfig mutators(GameMode::SOCCAR);
    if (g_lowGravity) {
        mutators.gravity = Vec(0, 0, -325.f);
        mutators.unlimitedFlips = true;
        mutators.demoMode = DemoMode::DISABLED;
    }
    Gym* gym = new Gym(match, TICK_SKIP, CAR_CONFIG_OCTANE, GameMode::SOCCAR, mutators);
    return {match, gym};
}

int main(int argc, char** argv) {
    const char* meshDir = argc > 1 ? argv[1] : "./collision_meshes";
    int iterations = argc > 2 ? atoi(argv[2]) : 5;
    for (int i = 1; i < argc; i++) {
        if (std::string(argv[i]) == "--custom-setter") g_customSetter = true;
        if (std::string(argv[i]) == "--low-gravity") g_lowGravity = true;
    }
    try {
        RocketSim::Init(meshDir);

        LearnerConfig cfg = {};
        cfg.numThreads = 16;
        cfg.numGamesPerThread = 24;
        int tsPerItr = 100 * 1000;
        cfg.timestepsPerIteration = tsPerItr;
        cfg.ppo.batchSize = tsPerItr;
        cfg.ppo.miniBatchSize = 25 * 1000;
        cfg.expBufferSize = tsPerItr * 3;
        cfg.ppo.epochs = 1;
        cfg.ppo.entCoef = 0.01f;
        cfg.ppo.policyLR = 2e-4f;
        cfg.ppo.criticLR = 2e-4f;
        cfg.ppo.policyLayerSizes = {256, 256, 256};
        cfg.ppo.criticLayerSizes = {256, 256, 256};

        ThreadAgentManager mgr(cfg);
        mgr.CreateAgents(EnvCreateFunc, cfg.numThreads, cfg.numGamesPerThread);

        // torch-default Linear init, U(+-1/sqrt(in)), for both networks
        std::mt19937 rng(cfg.randomSeed);
        const int obsSize = mgr.engine->ObsSize();
        for (int net = 0; net < 2; net++) {
            int in = obsSize;
            const IList& hidden = net == 0 ? cfg.ppo.policyLayerSizes : cfg.ppo.criticLayerSizes;
            for (size_t l = 0; l <= hidden.size(); l++) {
                int out = l < hidden.size() ? hidden[l] : (net == 0 ? RLG_NUM_ACTIONS : 1);
                std::uniform_real_distribution<float> u(-1.f / std::sqrt((float)in), 1.f / std::sqrt((float)in));
                std::vector<float> W((size_t)out * in), b(out);
                for (auto& x : W) x = u(rng);
                for (auto& x : b) x = u(rng);
                mgr.SetLayer(net, (int)l, W.data(), b.data(), out, in);
                in = out;
            }
        }
        mgr.StartAgents();

        uint64_t total = 0;
        double totalTime = 0, rewardSum = 0;
        for (int it = 0; it < iterations; it++) {
            auto t0 = std::chrono::steady_clock::now();
            rlg_traj_view v = mgr.CollectTimesteps(cfg.timestepsPerIteration);
            double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            uint64_t n = (uint64_t)v.T * v.N;
            std::vector<float> rew(n), adv(n);
            RLGB200::Check(rlg_engine_copy_to_host(mgr.engine->h, rew.data(), v.reward, n * 4));
            RLGB200::Check(rlg_engine_copy_to_host(mgr.engine->h, adv.data(), v.advantage, n * 4));
            double r = 0, a = 0;
            for (uint64_t i = 0; i < n; i++) { r += rew[i]; a += std::fabs(adv[i]); }
            Report rep;
            mgr.GetMetrics(rep);
            printf("iteration %d: Timesteps Collected %llu, Collected Steps/Second %.0f, Average Step Reward %.5f, Avg Advantage %.5f, Env Step Time %.4f, Policy Infer Time %.4f\n",
                   it, (unsigned long long)n, n / dt, r / n, a / n, rep["Env Step Time"], rep["Policy Infer Time"]);
            if (it > 0) { total += n; totalTime += dt; }
            rewardSum += r / n;
        }
        mgr.StopAgents();
        printf("{\"example\": \"examplemain\", \"custom_setter\": %s, \"arenas\": %d, \"steps_per_second\": %.0f, \"mean_step_reward\": %.6f}\n",
               g_customSetter ? "true" : "false", mgr.engine->NumArenas(), totalTime > 0 ? total / totalTime : 0.0, rewardSum / iterations);
    } catch (std::exception& e) {
        fprintf(stderr, "FATAL: %s\n", e.what());
        return 1;
    }
    return 0;
}
