// host_plugins.cpp — the host-plugin path against the fused path, bit for bit.
//
// Two ThreadAgentManagers with the same seeds, weights and arenas collect the same number of steps:
//   A  built-in plugins only                       -> everything fused on the device
//   B  the SAME maths as user-defined subclasses   -> obs builder, reward graph and a terminal condition run on the host from the
//      exported GameStates (OBSBuilder / RewardFunction / TerminalCondition virtuals, one Match per arena), StepCallback installed
// and the trajectories (states, actions, log-probs, rewards, next states, dones, truncateds) must be identical bit patterns: the
// GameState the host sees is the one the fused kernels read, and both sides do the same IEEE arithmetic in the same order.
//
//   ./host_plugins collision_meshes [arenas] [steps] [team_size]
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <rlgym_b200_shim.hpp>

using namespace RLGPC;
using namespace RLGSC;

static int g_teamSize = 1;

// ---- user-defined plugins restating the stock ones through the public plugin interface only ---------------------------------
class MyObs : public OBSBuilder {  // the DefaultOBS layout (G/Utils/OBSBuilders/DefaultOBS.cpp:3-55)
public:
    static void AddPlayer(FList& o, const PlayerData& p, bool inv) {
        const PhysObj& ph = p.GetPhys(inv);
        o += ph.pos * Vec(1 / CommonValues::SIDE_WALL_X, 1 / CommonValues::BACK_WALL_Y, 1 / CommonValues::CEILING_Z);
        o += ph.rotMat.forward;
        o += ph.rotMat.up;
        o += ph.vel * (1 / CommonValues::CAR_MAX_SPEED);
        o += ph.angVel * (1 / CommonValues::CAR_MAX_ANG_VEL);
        o += {p.boostFraction, (float)p.carState.isOnGround, (float)p.hasFlip, (float)p.carState.isDemoed};
    }
    FList BuildOBS(const PlayerData& player, const GameState& state, const Action& prevAction) override {
        FList r;
        const bool inv = player.team == Team::ORANGE;
        const PhysObj& ball = state.GetBallPhys(inv);
        r += ball.pos * Vec(1 / CommonValues::SIDE_WALL_X, 1 / CommonValues::BACK_WALL_Y, 1 / CommonValues::CEILING_Z);
        r += ball.vel * (1 / CommonValues::CAR_MAX_SPEED);
        r += ball.angVel * (1 / CommonValues::CAR_MAX_ANG_VEL);
        for (int i = 0; i < Action::ELEM_AMOUNT; i++) r += prevAction[i];
        const auto& pads = state.GetBoostPads(inv);
        for (int i = 0; i < CommonValues::BOOST_LOCATIONS_AMOUNT; i++) r += (float)pads[i];
        AddPlayer(r, player, inv);
        FList mates, opps;
        for (auto& other : state.players) {
            if (other.carId == player.carId) continue;
            AddPlayer(other.team == player.team ? mates : opps, other, inv);
        }
        r += mates;
        r += opps;
        return r;
    }
};
class MyFaceBall : public RewardFunction {
public:
    float GetReward(const PlayerData& p, const GameState& s, const Action&) override { return p.carState.rotMat.forward.Dot((s.ball.pos - p.phys.pos).Normalized()); }
};
class MyVelToBall : public RewardFunction {
public:
    float GetReward(const PlayerData& p, const GameState& s, const Action&) override {
        return (s.ball.pos - p.phys.pos).Normalized().Dot(p.phys.vel / CommonValues::CAR_MAX_SPEED);
    }
};
class MyNoTouch : public TerminalCondition {  // NoTouchCondition through the user interface
public:
    int steps = 0, maxSteps;
    explicit MyNoTouch(int m) : maxSteps(m) {}
    void Reset(const GameState&) override { steps = 0; }
    bool IsTerminal(const GameState& s) override {
        for (auto& p : s.players) if (p.ballTouchedStep) { steps = 0; return false; }
        return ++steps >= maxSteps;
    }
};

class MyDiscrete : public ActionParser {  // DiscreteAction's table (G/Utils/ActionParsers/DiscreteAction.cpp:3-67) through the user interface
public:
    std::vector<Action> table;
    MyDiscrete() {
        for (float throttle : {-1.f, 0.f, 1.f})
            for (float steer : {-1.f, 0.f, 1.f})
                for (float boost : {0.f, 1.f})
                    for (float handbrake : {0.f, 1.f}) {
                        if (boost == 1 && throttle != 1) continue;
                        table.push_back(Action{throttle, steer, 0, steer, 0, 0, boost, handbrake});
                    }
        for (float pitch : {-1.f, 0.f, 1.f})
            for (float yaw : {-1.f, 0.f, 1.f})
                for (float roll : {-1.f, 0.f, 1.f})
                    for (float jump : {0.f, 1.f})
                        for (float boost : {0.f, 1.f}) {
                            if (jump == 1 && yaw != 0) continue;
                            if (pitch == roll && roll == jump && jump == 0) continue;
                            const bool handbrake = jump == 1 && (pitch != 0 || yaw != 0 || roll != 0);
                            table.push_back(Action{boost, yaw, pitch, yaw, roll, jump, boost, (float)handbrake});
                        }
    }
    ActionSet ParseActions(const IList& idx, const GameState&) override {
        ActionSet out;
        for (int i : idx) out.push_back(table[i]);
        return out;
    }
    int GetActionAmount() override { return (int)table.size(); }
};

static EnvCreateResult MakeEnv(bool user) {
    RewardFunction* face = user ? (RewardFunction*)new MyFaceBall() : new FaceBallReward();
    RewardFunction* vel = user ? (RewardFunction*)new MyVelToBall() : new VelocityPlayerToBallReward();
    RewardFunction* rewards = new CombinedReward({{face, 0.1f}, {vel, 0.5f}, {new VelocityBallToGoalReward(), 1.0f},
                                                  {new EventReward({.teamGoal = 1.f, .concede = -1.f, .touch = 0.05f, .boostPickup = 0.1f}), 50.f}}, true);
    if (g_teamSize > 1) rewards = new ZeroSumReward(rewards, 0.3f, 1.f, true);
    std::vector<TerminalCondition*> conds = {user ? (TerminalCondition*)new MyNoTouch(40) : new NoTouchCondition(40), new GoalScoreCondition()};
    Match* match = new Match(rewards, conds, user ? (OBSBuilder*)new MyObs() : new DefaultOBS(), user ? (ActionParser*)new MyDiscrete() : new DiscreteAction(),
                             new RandomState(true, true, true), g_teamSize, true);
    return {match, new Gym(match, 8)};
}

static long g_callbackSteps = 0;
static void OnStep(GameInst*, const Gym::StepResult& r, Report& m) {
    m.AccumAvg("speed", r.state.players[0].phys.vel.Length());
    __atomic_add_fetch(&g_callbackSteps, 1, __ATOMIC_RELAXED);
}

int main(int argc, char** argv) {
    const char* meshDir = argc > 1 ? argv[1] : "./collision_meshes";
    const int arenas = argc > 2 ? atoi(argv[2]) : 256, steps = argc > 3 ? atoi(argv[3]) : 96;
    g_teamSize = argc > 4 ? atoi(argv[4]) : 1;
    try {
        RocketSim::Init(meshDir);
        PPOLearnerConfig pc;
        pc.policyLayerSizes = {64, 64};
        pc.criticLayerSizes = {64, 64};
        const int P = 2 * g_teamSize, obs = 51 + 19 * P;
        PPOLearner ppo(obs, RLG_NUM_ACTIONS, pc, Device{0}, 1024, 5);
        std::vector<GameTrajectory> traj(2);
        std::vector<std::vector<uint8_t>> bytes[2];
        ThreadAgentManager* mgrs[2];
        for (int k = 0; k < 2; k++) {
            auto* mgr = new ThreadAgentManager(ppo.policy, nullptr, nullptr, false, false, false, (uint64_t)arenas * P * steps, Device{0});
            mgrs[k] = mgr;
            mgr->ppoCfg = pc; mgr->randomSeed = 77; mgr->hostThreads = 8;
            mgr->CreateAgents([k] { return MakeEnv(k == 1); }, 1, arenas);
            if (k == 1) mgr->SetStepCallback(OnStep);
            RLGB200::Check(rlg_ppo_push_weights(ppo.h, mgr->collector, rlg_engine_stream(mgr->engine->h)));
            mgr->StartAgents();
            GameTrajectory t = mgr->CollectTimesteps((uint64_t)arenas * P * steps);
            traj[k] = t;
            const size_t n = t.size;
            const void* ptrs[7] = {t.data.states, t.data.actions, t.data.logProbs, t.data.rewards, t.data.nextStates, t.data.dones, t.data.truncateds};
            const size_t sizes[7] = {n * obs * 4, n * 8, n * 4, n * 4, n * obs * 4, n * 4, n * 4};
            for (int i = 0; i < 7; i++) {
                std::vector<uint8_t> h(sizes[i]);
                RLGB200::Check(rlg_engine_copy_to_host(mgr->engine->h, h.data(), ptrs[i], sizes[i]));
                bytes[k].push_back(std::move(h));
            }
        }
        const char* names[7] = {"states", "actions", "logProbs", "rewards", "nextStates", "dones", "truncateds"};
        long mismatches = 0;
        for (int i = 0; i < 7; i++) {
            long bad = 0;
            for (size_t j = 0; j < bytes[0][i].size(); j += 4) bad += memcmp(&bytes[0][i][j], &bytes[1][i][j], 4) != 0;
            if (bad) fprintf(stderr, "%s: %ld of %zu words differ\n", names[i], bad, bytes[0][i].size() / 4);
            mismatches += bad;
        }
        double doneSum = 0, rewAbs = 0;
        const float* dn = (const float*)bytes[1][5].data();
        const float* rw = (const float*)bytes[1][3].data();
        for (size_t j = 0; j < traj[1].size; j++) { doneSum += dn[j]; rewAbs += std::fabs(rw[j]); }
        Report repA, repB;
        mgrs[0]->GetMetrics(repA);
        mgrs[1]->GetMetrics(repB);
        double speed = 0; long games = 0;
        for (auto* g : mgrs[1]->gameInsts) if (g->_metrics.Has("speed_avg_count")) { speed += g->_metrics.GetAvg("speed"); games++; }
        printf("{\"rows\": %zu, \"mismatched_words\": %ld, \"host_path_a\": %s, \"host_path_b\": %s, \"episodes_ended\": %.0f, \"mean_abs_reward\": %.6f, "
               "\"callback_steps\": %ld, \"avg_step_reward_fused\": %.9g, \"avg_step_reward_host\": %.9g, \"avg_episode_reward_fused\": %.9g, "
               "\"avg_episode_reward_host\": %.9g, \"mean_speed\": %.3f}\n",
               traj[1].size, mismatches, mgrs[0]->hostPath ? "true" : "false", mgrs[1]->hostPath ? "true" : "false", doneSum, rewAbs / traj[1].size, g_callbackSteps,
               repA["Average Step Reward"], repB["Average Step Reward"], repA["Average Episode Reward"], repB["Average Episode Reward"], games ? speed / games : 0.0);
        delete mgrs[0];
        delete mgrs[1];
        return mismatches == 0 ? 0 : 2;
    } catch (std::exception& e) {
        fprintf(stderr, "FATAL: %s\n", e.what());
        return 1;
    }
}
