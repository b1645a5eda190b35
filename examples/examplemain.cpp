// examplemain.cpp — a training app written against the reference's public surface (RLGPC::Learner, EnvCreateFn, LearnerConfig,
// step / iteration callbacks; compare T/examplemain.cpp) running on the B200 engine: collection AND the PPO update happen on the
// GPU (csrc/*.cu), this file is configuration.  The reference's own examplemain.cpp also compiles unmodified against
// include/compat (tests/test_cpp_shim.py); this variant adds command-line switches for the tests:
//
//   g++ -std=c++20 -O2 -Iinclude -Iinclude/compat examples/examplemain.cpp -o examplemain -lpthread
//       -Lrlgymppo_cpp_b200/csrc -lrlgym_b200 -Wl,-rpath,$PWD/rlgymppo_cpp_b200/csrc   (one line)
//   python -m rlgymppo_cpp_b200.meshes --out collision_meshes
//   ./examplemain collision_meshes [iterations] [--custom-setter] [--low-gravity] [--step-callback] [--user-reward] [--save DIR]
#include <cstdio>
#include <cstdlib>
#include <random>

#include <RLGymPPO_CPP/Learner.h>
#include <RLGymSim_CPP/Utils/OBSBuilders/DefaultOBS.h>
#include <RLGymSim_CPP/Utils/RewardFunctions/CombinedReward.h>
#include <RLGymSim_CPP/Utils/RewardFunctions/CommonRewards.h>

using namespace RLGPC;  // RLGymPPO
using namespace RLGSC;  // RLGymSim

static bool g_customSetter = false, g_lowGravity = false, g_stepCallback = false, g_userReward = false;

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

// A user-defined RewardFunction (host-plugin path): closer to the ball is better, touching it is best.
class BallProximityReward : public RewardFunction {
public:
    float GetReward(const PlayerData& player, const GameState& state, const Action&) override {
        const float dist = player.phys.pos.Dist(state.ball.pos);
        return (player.ballTouchedStep ? 1.f : 0.f) + std::exp(-dist / 2000.f);
    }
};

// Step callback (GameInst.h:7): per-game metrics, gathered by the iteration callback below
static void OnStep(GameInst*, const Gym::StepResult& stepResult, Report& gameMetrics) {
    for (auto& player : stepResult.state.players) {
        gameMetrics.AccumAvg("player_speed", player.phys.vel.Length());
        gameMetrics.AccumAvg("ball_touch_ratio", player.ballTouchedStep);
        gameMetrics.AccumAvg("in_air_ratio", !player.carState.isOnGround);
    }
}
static double g_lastSpeed = -1, g_lastAir = -1;
static void OnIteration(Learner* learner, Report& allMetrics) {
    AvgTracker speed, air;
    for (auto& r : learner->GetAllGameMetrics()) {
        if (!r.Has("player_speed_avg_count")) continue;
        speed += (float)r.GetAvg("player_speed");
        air += (float)r.GetAvg("in_air_ratio");
    }
    if (speed.count) { allMetrics["player_speed"] = g_lastSpeed = speed.Get(); allMetrics["in_air_ratio"] = g_lastAir = air.Get(); }
}

EnvCreateResult EnvCreateFunc() {
    constexpr int TICK_SKIP = 8;
    constexpr float NO_TOUCH_TIMEOUT_SECS = 10.f;
    std::vector<std::pair<RewardFunction*, float>> terms = {
        {new FaceBallReward(), 0.1f},
        {new VelocityPlayerToBallReward(), 0.5f},
        {new VelocityBallToGoalReward(), 1.0f},
        {new EventReward({.teamGoal = 1.f, .concede = -1.f}), 50.f},
    };
    if (g_userReward) terms.push_back({new BallProximityReward(), 0.25f});  // a stock / user mix: the reward graph moves to the host
    auto rewards = new CombinedReward(terms, true);
    std::vector<TerminalCondition*> terminalConditions = {new NoTouchCondition(NO_TOUCH_TIMEOUT_SECS * 120 / TICK_SKIP), new GoalScoreCondition()};
    StateSetter* stateSetter = g_customSetter ? (StateSetter*)new MidfieldDropState() : (StateSetter*)new RandomState(true, true, true);
    Match* match = new Match(rewards, terminalConditions, new DefaultOBS(), new DiscreteAction(), stateSetter, 1, true);
    MutatorConfig mutators(GameMode::SOCCAR);
    if (g_lowGravity) {
        mutators.gravity = Vec(0, 0, -325.f);
        mutators.unlimitedFlips = true;
        mutators.demoMode = DemoMode::DISABLED;
    }
    return {match, new Gym(match, TICK_SKIP, CAR_CONFIG_OCTANE, GameMode::SOCCAR, mutators)};
}

int main(int argc, char** argv) {
    const char* meshDir = argc > 1 ? argv[1] : "./collision_meshes";
    const int iterations = argc > 2 ? atoi(argv[2]) : 5;
    std::string saveDir;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "--custom-setter") g_customSetter = true;
        if (a == "--low-gravity") g_lowGravity = true;
        if (a == "--step-callback") g_stepCallback = true;
        if (a == "--user-reward") g_userReward = true;
        if (a == "--save" && i + 1 < argc) saveDir = argv[i + 1];
    }
    try {
        RocketSim::Init(meshDir);

        LearnerConfig cfg = {};
        cfg.numThreads = 16;
        cfg.numGamesPerThread = 24;
        const int tsPerItr = 100 * 1000;
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
        cfg.sendMetrics = false;
        cfg.checkpointLoadFolder = saveDir;
        cfg.checkpointSaveFolder = saveDir;
        cfg.timestepsPerSave = 150 * 1000;

        Learner learner(EnvCreateFunc, cfg);
        learner.maxIterations = (uint64_t)iterations;
        if (g_stepCallback) learner.stepCallback = OnStep;
        double firstEntropy = -1, lastEntropy = -1, stepRew = 0, sps = 0;
        uint64_t updates = 0, collected = 0;
        int seen = 0;
        learner.iterationCallback = [&](Learner* l, Report& rep) {
            OnIteration(l, rep);
            if (firstEntropy < 0) firstEntropy = rep["Policy Entropy"];
            lastEntropy = rep["Policy Entropy"];
            updates = (uint64_t)rep["Cumulative Model Updates"];
            collected = (uint64_t)rep["Timesteps Collected"];
            stepRew += rep["Average Step Reward"];
            if (seen > 0) sps += rep["Collected Steps/Second"];
            seen++;
        };
        const uint64_t startTimesteps = learner.totalTimesteps;
        learner.Learn();
        printf("{\"example\": \"examplemain\", \"custom_setter\": %s, \"host_path\": %s, \"arenas\": %d, \"iterations\": %d, \"timesteps_collected\": %llu, "
               "\"start_timesteps\": %llu, \"total_timesteps\": %llu, \"model_updates\": %llu, \"first_entropy\": %.6f, \"last_entropy\": %.6f, "
               "\"mean_step_reward\": %.6f, \"steps_per_second\": %.0f, \"player_speed\": %.3f, \"in_air_ratio\": %.4f}\n",
               g_customSetter ? "true" : "false", learner.agentMgr->hostPath ? "true" : "false", learner.agentMgr->engine->NumArenas(), seen,
               (unsigned long long)collected, (unsigned long long)startTimesteps, (unsigned long long)learner.totalTimesteps, (unsigned long long)updates,
               firstEntropy, lastEntropy, seen ? stepRew / seen : 0.0, seen > 1 ? sps / (seen - 1) : 0.0, g_lastSpeed, g_lastAir);
    } catch (std::exception& e) {
        fprintf(stderr, "FATAL: %s\n", e.what());
        return 1;
    }
    return 0;
}
