// rl_gym.h — the RLGymSim "gym" layer as straight-line device code:
// action parse, GameEventTracker, GameState snapshot, DefaultOBS / DefaultOBSPadded,
// CombinedReward / EventReward / ZeroSumReward, terminal conditions, state setters.
//
// Everything here must be BIT-EXACT against the reference given identical arena
// states, so each expression keeps the reference's operation order and every floating-point
// operation whose result feeds another one goes through the strict s_* helpers of rl_math.h
// (the oracle is built for baseline x86-64: no FMA, IEEE division and square root), while the
// physics around it is free to use FMA contraction.
// Reference paths: G/ = RLGymPPO_CPP/RLGymSim_CPP/src/RLGymSim_CPP, R/ = .../RocketSim/src.
#pragma once
#include "rl_car.h"

namespace rl {

struct Tables {
    float actions[96 * 8];    // DiscreteAction table (G/Utils/ActionParsers/DiscreteAction.cpp:3-67: 90 rows) or a user parser's (SimCfg::numActions rows)
    int32_t padMap[kNumPads]; // GameState pad i -> RocketSim pad index (G/Utils/Gamestates/GameState.cpp:10-50)
    float padPos[kNumPads * 3]; // RocketSim pad order (6 big, 28 small), uu
    float padPosBT[kNumPads * 3]; // same, Bullet units (uu * (1/50), as BoostPad::_BulletSetup stores it)
    // BoostPadGrid (BoostPadGrid.cpp:5-25): for car cell (ix, iy), ix in [-1, 8], iy in [-1, 10], the pads whose own cell
    // lies in the clamped 3x3 neighbourhood; entry (ix+1) + 10*(iy+1), lo/hi words of a 34-bit mask
    uint32_t padCellMask[10 * 12 * 2];
};

RL_HDI V3 car_pos_uu(const CarS& c) { return to_uu(c.pos); }
RL_HDI V3 car_vel_uu(const CarS& c) { return to_uu(c.vel); }
RL_HDI V3 ball_pos_uu(const BallS& b) { return to_uu(b.pos); }
RL_HDI V3 ball_vel_uu(const BallS& b) { return to_uu(b.vel); }

// G/Math.cpp:3-5
RL_HDI bool is_ball_scored_y(float yUU) { return fabsf(yUU) > C::GOAL_THRESHOLD_Y + C::BALL_RADIUS; }

// ---- action parse: Match::ParseActions + Gym::Step:69-79 --------------------------------
RL_HDI void parse_actions(ArenaS& a, const SimCfg& cfg, const Tables& tb, const int32_t* actionIdx /*player order*/) {
    for (int p = 0; p < cfg.numCars; p++) {
        CarS& car = a.cars[cfg.playerOrder[p]];
        float act[8];
        bool zero = car.snapIsDemoed != 0;  // Match.cpp:47-49 reads the PREVIOUS snapshot's carState
        int idx = actionIdx[p];
        for (int k = 0; k < 8; k++) act[k] = zero ? 0.f : tb.actions[idx * 8 + k];
        for (int k = 0; k < 8; k++) car.prevAction[k] = act[k];
        // (CarControls)Action, G/Utils/BasicTypes/Action.h:34-47
        car.controls.throttle = act[0]; car.controls.steer = act[1];
        car.controls.pitch = act[2]; car.controls.yaw = act[3]; car.controls.roll = act[4];
        car.controls.jump = act[5] == 1.f; car.controls.boost = act[6] == 1.f; car.controls.handbrake = act[7] == 1.f;
    }
}

// ---- GameEventTracker (R/Sim/GameEventTracker/GameEventTracker.cpp) -----------------------
// Arena::IsBallProbablyGoingIn, soccar branch (R/Sim/Arena/Arena.cpp:827-863)
RL_HD RL_NOINLINE inline bool ball_probably_going_in(const BallS& b, const Mut& mu, float maxTime, float extraMargin, int* goalTeamOut) {
    V3 pos = ball_pos_uu(b), vel = ball_vel_uu(b);
    if (fabsf(vel.y) < kEps) return false;
    float scoreDirSgn = (float)sgn(vel.y);
    float goalY = mu.goalBaseThresholdY * scoreDirSgn;
    float distToGoal = fabsf(s_sub(pos.y, goalY));
    float timeToGoal = s_div(distToGoal, fabsf(vel.y));
    if (timeToGoal > maxTime) return false;
    // ballPos + ballVel*t + gravity*t*t/2 ; gravity = (0,0,-650)
    float ex = s_add(s_add(pos.x, s_mul(vel.x, timeToGoal)), s_div(s_mul(s_mul(mu.gravityX, timeToGoal), timeToGoal), 2.f));
    float ez = s_add(s_add(pos.z, s_mul(vel.z, timeToGoal)), s_div(s_mul(s_mul(mu.gravityZ, timeToGoal), timeToGoal), 2.f));
    const float APPROX_GOAL_HALF_WIDTH = 892.755f, APPROX_GOAL_HEIGHT = (float)642.775;
    float scoreMargin = s_add(s_mul(mu.ballRadius, 0.1f), extraMargin);
    if (ez > APPROX_GOAL_HEIGHT + scoreMargin) return false;
    if (fabsf(ex) > APPROX_GOAL_HALF_WIDTH + scoreMargin) return false;
    if (goalTeamOut) *goalTeamOut = scoreDirSgn < 0 ? 0 : 1;  // RS_TEAM_FROM_Y(scoreDirSgn)
    return true;
}

// GetShooterPasser (GameEventTracker.cpp:5-46); returns car indices or -1. Iterates in _cars order.
RL_HD RL_NOINLINE inline bool get_shooter_passer(const ArenaS& a, const SimCfg& cfg, int team, int& shooter, bool findPasser, int& passer,
                               int64_t maxShooterTicks, int64_t maxPasserTicks) {
    shooter = passer = -1;
    int64_t tick = get_i64(a.tickLo, a.tickHi);
    int64_t shooterHit = 0, passerHit = 0;
    for (int p = 0; p < cfg.numCars; p++) {
        int ci = cfg.playerOrder[p];
        const CarS& c = a.cars[ci];
        if (car_team(ci, cfg.spawnOpponents) != team || !c.hitValid) continue;
        int64_t hit = get_i64(c.hitTickLo, c.hitTickHi);
        if (hit + maxShooterTicks >= tick) {
            if (shooter < 0 || hit > shooterHit) { shooter = ci; shooterHit = hit; }
        }
    }
    if (shooter >= 0 && findPasser) {
        for (int p = 0; p < cfg.numCars; p++) {
            int ci = cfg.playerOrder[p];
            const CarS& c = a.cars[ci];
            if (car_team(ci, cfg.spawnOpponents) != team || !c.hitValid || ci == shooter) continue;
            int64_t hit = get_i64(c.hitTickLo, c.hitTickHi);
            if (hit + maxPasserTicks >= shooterHit) {
                if (passer < 0 || hit > passerHit) { passer = ci; passerHit = hit; }
            }
        }
    }
    return shooter >= 0;
}

// GameEventTracker::Update with the Gym's callbacks (G/Gym.cpp:6-38) folded in.
RL_HDI void event_tracker_update(ArenaS& a, const SimCfg& cfg) {
    // default GameEventTrackerConfig (GameEventTracker.h:11-40); tick rate 120
    const float shotMinSpeed = 1750, predScoreExtraMargin = 0, shotEventCooldown = 1.0f, shotMinScoreTime = 2.0f;
    const int64_t goalMaxTouchTicks = 480, passMaxTouchTicks = 240, shotMinTouchDelayTicks = 36;
    bool scored = fabsf(s_mul(a.ball.pos.y, BT2UU)) > cfg.mut.goalBaseThresholdY + cfg.mut.ballRadius;  // Arena::IsBallScored (Arena.cpp:949-957)
    int32_t cnt = a.ball.updateCounterLo;
    if (cnt > a.lastBallUpdateCount) {
        int64_t deltaTicks = (int64_t)cnt - a.lastBallUpdateCount;
        float deltaTime = s_mul((float)deltaTicks, kTickTime);
        if (scored && !a.ballScoredLast) {
            int shooter, passer;
            int team = (-a.ball.pos.y) < 0 ? 0 : 1;  // RS_TEAM_FROM_Y(-ball y)
            if (get_shooter_passer(a, cfg, team, shooter, true, passer, goalMaxTouchTicks, passMaxTouchTicks)) {
                a.cars[shooter].matchGoals++;
                if (passer >= 0) a.cars[passer].matchAssists++;
            }
        } else if (!a.ballShot) {
            if (a.shotCooldown > 0) {
                a.shotCooldown = fmaxf_(s_sub(a.shotCooldown, deltaTime), 0.f);
            } else {
                V3 v = ball_vel_uu(a.ball);
                float speedSq = s_add(s_add(s_mul(v.x, v.x), s_mul(v.y, v.y)), s_mul(v.z, v.z));  // Vec::LengthSq via btVector3::length2
                if (speedSq >= shotMinSpeed * shotMinSpeed) {
                    int goalTeam = 0;
                    if (ball_probably_going_in(a.ball, cfg.mut, shotMinScoreTime, predScoreExtraMargin, &goalTeam)) {
                        int shooterTeam = 1 - goalTeam;
                        int shooter, passer;
                        if (get_shooter_passer(a, cfg, shooterTeam, shooter, true, passer,
                                               deltaTicks + shotMinTouchDelayTicks, passMaxTouchTicks)) {
                            int64_t since = get_i64(a.tickLo, a.tickHi) - get_i64(a.cars[shooter].hitTickLo, a.cars[shooter].hitTickHi);
                            if (since >= shotMinTouchDelayTicks) {
                                a.ballShot = 1;
                                a.ballShotGoalTeam = goalTeam;
                                a.shotCooldown = shotEventCooldown;
                                a.cars[shooter].matchShots++;
                                if (passer >= 0) a.cars[passer].matchShotPasses++;
                            }
                        }
                    }
                }
            }
        } else {
            bool willScore = ball_probably_going_in(a.ball, cfg.mut, shotMinScoreTime, predScoreExtraMargin, nullptr);
            if (!willScore) {
                int saver, unused;
                if (get_shooter_passer(a, cfg, a.ballShotGoalTeam, saver, false, unused, deltaTicks, 0)) a.cars[saver].matchSaves++;
                a.ballShot = 0;
            }
        }
    } else if (cnt == a.lastBallUpdateCount) {
        return;
    } else {
        a.ballScoredLast = 0; a.ballShot = 0; a.shotCooldown = 0;  // ResetPersistentInfo
    }
    a.ballScoredLast = scored;
    a.lastBallUpdateCount = cnt;
}

// ---- GameState::UpdateFromArena (G/Utils/Gamestates/GameState.cpp:52-104) ------------------
RL_HDI void snapshot_update(ArenaS& a, const SimCfg& cfg) {
    int64_t tick = get_i64(a.tickLo, a.tickHi);
    int64_t last = get_i64(a.lastTickLo, a.lastTickHi);
    int64_t tickSkip = tick - last; if (tickSkip < 0) tickSkip = 0;
    for (int p = 0; p < cfg.numCars; p++) {
        int ci = cfg.playerOrder[p];
        CarS& c = a.cars[ci];
        // PlayerData::UpdateFromCar (PlayerData.cpp:20-25)
        c.touchedStep = c.hitValid ? (get_i64(c.hitTickLo, c.hitTickHi) >= (tick - tickSkip)) : 0;
        if (c.touchedStep) a.lastTouchCarId = ci + 1;
        c.snapIsDemoed = c.isDemoed;
    }
    float by = s_mul(a.ball.pos.y, BT2UU);
    if (is_ball_scored_y(by)) a.scoreLine[1 - (by < 0 ? 0 : 1)]++;
    set_i64(a.lastTickLo, a.lastTickHi, tick);
}

// ---- DefaultOBS / DefaultOBSPadded -----------------------------------------------------------
RL_HDI int add_player_obs(float* o, const CarS& c, bool inv) {
    const float px = 1 / C::ARENA_EXTENT_X, py = 1 / C::ARENA_EXTENT_Y, pz = 1 / 2044.f;
    const float velCoef = 1 / C::CAR_MAX_SPEED, angCoef = 1 / C::CAR_MAX_ANG_SPEED;
    V3 pos = car_pos_uu(c), vel = car_vel_uu(c), ang = c.angvel;
    V3 fwd = c.rot.col(0), up = c.rot.col(2);
    if (inv) {  // PhysObj::Invert (PhysObj.cpp:19-31): multiply by (-1,-1,1)
        pos = V3(pos.x * -1.f, pos.y * -1.f, pos.z * 1.f); vel = V3(vel.x * -1.f, vel.y * -1.f, vel.z * 1.f);
        ang = V3(ang.x * -1.f, ang.y * -1.f, ang.z * 1.f);
        fwd = V3(fwd.x * -1.f, fwd.y * -1.f, fwd.z * 1.f); up = V3(up.x * -1.f, up.y * -1.f, up.z * 1.f);
    }
    o[0] = pos.x * px; o[1] = pos.y * py; o[2] = pos.z * pz;
    o[3] = fwd.x; o[4] = fwd.y; o[5] = fwd.z;
    o[6] = up.x; o[7] = up.y; o[8] = up.z;
    o[9] = vel.x * velCoef; o[10] = vel.y * velCoef; o[11] = vel.z * velCoef;
    o[12] = ang.x * angCoef; o[13] = ang.y * angCoef; o[14] = ang.z * angCoef;
    o[15] = s_div(c.boost, 100.f);         // PlayerData.cpp:32
    o[16] = (float)(c.isOnGround != 0);
    bool hasFlip = !c.hasDoubleJumped && !c.hasFlipped && c.airTimeSinceJump < C::DOUBLEJUMP_MAX_DELAY;  // PlayerData.cpp:28-30
    o[17] = (float)hasFlip;
    o[18] = (float)(c.isDemoed != 0);
    return 19;
}

// Fisher-Yates with the arena RNG (the reference uses std::shuffle with the thread RNG,
// DefaultOBSPadded.cpp:58-59; slot order is random on both sides, only the multiset is comparable)
RL_HDI void shuffle_slots(ArenaS& a, int* slots, int n) {
    for (int i = n - 1; i > 0; i--) {
        int j = (int)(rng_next(a) % (uint32_t)(i + 1));
        int t = slots[i]; slots[i] = slots[j]; slots[j] = t;
    }
}

// obs row stride is `stride` floats (so a transposed/strided output buffer can be used)
RL_HD RL_NOINLINE inline void build_obs(ArenaS& a, const SimCfg& cfg, const Tables& tb, float* out /*[P][obsSize]*/) {
    const float px = 1 / C::ARENA_EXTENT_X, py = 1 / C::ARENA_EXTENT_Y, pz = 1 / 2044.f;
    const float velCoef = 1 / C::CAR_MAX_SPEED, angCoef = 1 / C::CAR_MAX_ANG_SPEED;
    for (int p = 0; p < cfg.numCars; p++) {
        int ci = cfg.playerOrder[p];
        const CarS& me = a.cars[ci];
        int team = car_team(ci, cfg.spawnOpponents);
        bool inv = team == 1;
        float* o = out + (size_t)p * cfg.obsSize;
        V3 bp = ball_pos_uu(a.ball), bv = ball_vel_uu(a.ball), ba = a.ball.angvel;
        if (inv) {
            bp = V3(bp.x * -1.f, bp.y * -1.f, bp.z * 1.f); bv = V3(bv.x * -1.f, bv.y * -1.f, bv.z * 1.f);
            ba = V3(ba.x * -1.f, ba.y * -1.f, ba.z * 1.f);
        }
        int k = 0;
        o[k++] = bp.x * px; o[k++] = bp.y * py; o[k++] = bp.z * pz;
        o[k++] = bv.x * velCoef; o[k++] = bv.y * velCoef; o[k++] = bv.z * velCoef;
        o[k++] = ba.x * angCoef; o[k++] = ba.y * angCoef; o[k++] = ba.z * angCoef;
        for (int i = 0; i < 8; i++) o[k++] = me.prevAction[i];
        for (int i = 0; i < kNumPads; i++) {
            int gi = inv ? (kNumPads - i - 1) : i;  // GameState.cpp:84-91
            o[k++] = (float)((pads_active(a.pads) >> tb.padMap[gi]) & 1ULL);
        }
        k += add_player_obs(o + k, me, inv);
        if (cfg.obsKind == 0) {
            // teammates then opponents, each in _cars order (DefaultOBS.cpp:40-53)
            for (int pass = 0; pass < 2; pass++)
                for (int q = 0; q < cfg.numCars; q++) {
                    int cj = cfg.playerOrder[q];
                    if (cj == ci) continue;
                    bool mate = car_team(cj, cfg.spawnOpponents) == team;
                    if (mate == (pass == 0)) k += add_player_obs(o + k, a.cars[cj], inv);
                }
        } else {
            int mp = cfg.obsMaxPlayers;
            int mates[kMaxCars], opps[kMaxCars];
            int nm = 0, no = 0;
            for (int q = 0; q < cfg.numCars; q++) {
                int cj = cfg.playerOrder[q];
                if (cj == ci) continue;
                if (car_team(cj, cfg.spawnOpponents) == team) mates[nm++] = cj; else opps[no++] = cj;
            }
            while (nm < mp - 1) mates[nm++] = -1;
            while (no < mp) opps[no++] = -1;
            shuffle_slots(a, mates, nm);
            shuffle_slots(a, opps, no);
            for (int i = 0; i < nm; i++) {
                if (mates[i] >= 0) k += add_player_obs(o + k, a.cars[mates[i]], inv);
                else { for (int z = 0; z < 19; z++) o[k + z] = 0.f; k += 19; }
            }
            for (int i = 0; i < no; i++) {
                if (opps[i] >= 0) k += add_player_obs(o + k, a.cars[opps[i]], inv);
                else { for (int z = 0; z < 19; z++) o[k + z] = 0.f; k += 19; }
            }
        }
    }
}

// ---- rewards ------------------------------------------------------------------------------
RL_HDI void event_values(const ArenaS& a, const SimCfg& cfg, int ci, float* v) {
    const CarS& c = a.cars[ci];
    int team = car_team(ci, cfg.spawnOpponents);
    v[0] = (float)c.matchGoals; v[1] = (float)a.scoreLine[team]; v[2] = (float)a.scoreLine[1 - team];
    v[3] = (float)c.matchAssists; v[4] = (float)(c.touchedStep != 0); v[5] = (float)c.matchShots;
    v[6] = (float)c.matchShotPasses; v[7] = (float)c.matchSaves; v[8] = (float)c.matchDemos;
    v[9] = (float)(c.isDemoed != 0); v[10] = s_div(c.boost, 100.f);
}

RL_HDI float reward_term(ArenaS& a, const SimCfg& cfg, const RewardTerm& t, int ci) {
    CarS& c = a.cars[ci];
    V3 ballPos = ball_pos_uu(a.ball);
    switch (t.kind) {
    case 0: {  // EventReward::GetReward (CommonRewards.cpp:32-43)
        float nv[11]; event_values(a, cfg, ci, nv);
        float r = 0;
        for (int i = 0; i < 11; i++) { r = s_add(r, s_mul(fmaxf_(s_sub(nv[i], c.eventMemo[i]), 0.f), t.params[i])); c.eventMemo[i] = nv[i]; }
        return r;
    }
    case 1: {  // VelocityPlayerToBallReward (CommonRewards.h:91-98)
        V3 d = ref_normalized(s_sub3(ballPos, car_pos_uu(c)));
        V3 v = car_vel_uu(c);
        V3 nvv = V3(s_div(v.x, C::CAR_MAX_SPEED), s_div(v.y, C::CAR_MAX_SPEED), s_div(v.z, C::CAR_MAX_SPEED));
        return ref_dot(d, nvv);
    }
    case 2: {  // VelocityBallToGoalReward (CommonRewards.h:73-88)
        bool targetOrange = car_team(ci, cfg.spawnOpponents) == 0;
        if (t.params[0] != 0.f) targetOrange = !targetOrange;
        const float goalZ = ((float)642.775) / 2;
        V3 target = targetOrange ? V3(0, 6000, goalZ) : V3(0, -6000, goalZ);
        V3 d = ref_normalized(s_sub3(target, ballPos));
        V3 v = ball_vel_uu(a.ball);
        V3 nvv = V3(s_div(v.x, C::BALL_MAX_SPEED), s_div(v.y, C::BALL_MAX_SPEED), s_div(v.z, C::BALL_MAX_SPEED));
        return ref_dot(d, nvv);
    }
    case 3: {  // FaceBallReward (CommonRewards.h:101-108)
        V3 d = ref_normalized(s_sub3(ballPos, car_pos_uu(c)));
        return ref_dot(c.rot.col(0), d);
    }
    case 4: {  // VelocityReward (CommonRewards.h:52-58)
        float neg = t.params[0] != 0.f ? 1.f : 0.f;
        return s_mul(s_div(ref_len(car_vel_uu(c)), C::CAR_MAX_SPEED), (float)(1 - 2 * (int)neg));
    }
    // the two powf rewards: glibc's powf evaluates in double and rounds once; so does this (within 1 ulp of it, tests say so)
    case 5: {  // SaveBoostReward (CommonRewards.h:61-70): RS_CLAMP(powf(boostFraction, exponent), 0, 1), boostFraction = boost / 100
        float frac = s_div(c.boost, 100.f);
        float v = (float)pow((double)frac, (double)t.params[0]);
        return fminf_(fmaxf_(v, 0.f), 1.f);
    }
    case 6: {  // TouchBallReward (CommonRewards.h:110-124)
        if (!c.touchedStep) return 0.f;
        float x = s_div(s_add(ballPos.z, C::BALL_RADIUS), s_mul(C::BALL_RADIUS, 2.f));
        return (float)pow((double)x, (double)t.params[0]);
    }
    }
    return 0.f;
}

// CombinedReward::GetAllRewards (+ ZeroSumReward::GetAllRewards), output in player order
RL_HD RL_NOINLINE inline void compute_rewards(ArenaS& a, const SimCfg& cfg, float* out /*[P]*/) {
    float r[kMaxCars];
    for (int p = 0; p < cfg.numCars; p++) r[p] = 0.f;
    for (int i = 0; i < cfg.numRewardTerms; i++)
        for (int p = 0; p < cfg.numCars; p++) r[p] = s_add(r[p], s_mul(reward_term(a, cfg, cfg.rewards[i], cfg.playerOrder[p]), cfg.rewards[i].weight));
    if (cfg.zeroSum) {  // ZeroSumReward.cpp:3-29
        int cnt[2] = {0, 0};
        float avg[2] = {0.f, 0.f};
        for (int p = 0; p < cfg.numCars; p++) { int t = car_team(cfg.playerOrder[p], cfg.spawnOpponents); cnt[t]++; avg[t] = s_add(avg[t], r[p]); }
        for (int t = 0; t < 2; t++) avg[t] = s_div(avg[t], (float)(cnt[t] > 1 ? cnt[t] : 1));
        for (int p = 0; p < cfg.numCars; p++) {
            int t = car_team(cfg.playerOrder[p], cfg.spawnOpponents);
            r[p] = s_sub(s_add(s_mul(r[p], s_sub(1.f, cfg.teamSpirit)), s_mul(avg[t], cfg.teamSpirit)), s_mul(avg[1 - t], cfg.opponentScale));
        }
    }
    for (int p = 0; p < cfg.numCars; p++) out[p] = r[p];
}

// Match::IsDone with [NoTouchCondition, GoalScoreCondition] (examplemain.cpp:78-81 order)
RL_HDI bool compute_done(ArenaS& a, const SimCfg& cfg) {
    bool done = false;
    if (cfg.noTouchMaxSteps > 0) {  // NoTouchCondition.h:18-28
        bool touched = false;
        for (int c = 0; c < cfg.numCars; c++) touched |= a.cars[c].touchedStep != 0;
        if (touched) a.stepsSinceTouch = 0;
        else { a.stepsSinceTouch++; done = a.stepsSinceTouch >= cfg.noTouchMaxSteps; }
    }
    if (!done && cfg.goalScoreTerminal) done = is_ball_scored_y(s_mul(a.ball.pos.y, BT2UU));
    return done;
}

// Match::EpisodeReset + Gym::Reset bookkeeping after the arena has been set to its new state
RL_HDI void episode_reset(ArenaS& a, const SimCfg& cfg) {
    // GameState(arena): fresh counters, scoreLine, lastTickCount = now (GameState.cpp:52-104)
    a.scoreLine[0] = a.scoreLine[1] = 0;
    a.lastTouchCarId = -1;
    a.lastTickLo = a.lastTickHi = 0;
    for (int c = 0; c < cfg.numCars; c++) {
        CarS& car = a.cars[c];
        car.matchGoals = car.matchSaves = car.matchAssists = car.matchShots = 0;
        car.matchShotPasses = car.matchBumps = car.matchDemos = car.boostPickups = 0;
        for (int k = 0; k < 8; k++) car.prevAction[k] = 0.f;
    }
    snapshot_update(a, cfg);
    a.stepsSinceTouch = 0;
    for (int c = 0; c < cfg.numCars; c++) event_values(a, cfg, c, a.cars[c].eventMemo);  // EventReward::Reset
    a.ballScoredLast = 0; a.ballShot = 0; a.shotCooldown = 0;  // eventTracker.ResetPersistentInfo()
}

// ---- state setters ------------------------------------------------------------------------
RL_HDI void car_set_pose(CarS& c, V3 posUU, float yaw, float pitch, float roll) {
    c.pos = V3(posUU.x * UU2BT, posUU.y * UU2BT, posUU.z * UU2BT);
    c.rot = angle_to_rotmat(yaw, pitch, roll);
}

// Arena::ResetToRandomKickoff (Arena.cpp:112-216) + Match::ResetState pad reset (Match.cpp:66-67)
RL_HDI void reset_to_kickoff(ArenaS& a, const SimCfg& cfg) {
    const float SX[5] = {-2048, 2048, -256, 256, 0};
    const float SY[5] = {-2560, -2560, -3840, -3840, -4608};
    const double PI4 = 0.78539816339744830962;
    const float SYAW[5] = {(float)(PI4 * 1), (float)(PI4 * 3), (float)(PI4 * 2), (float)(PI4 * 2), (float)(PI4 * 2)};
    int order[5] = {0, 1, 2, 3, 4};
    shuffle_slots(a, order, 5);
    int nTeam[2] = {0, 0};
    for (int ci = 0; ci < cfg.numCars; ci++) {  // cars of a team in _cars order
        int c = cfg.playerOrder[ci];
        int team = car_team(c, cfg.spawnOpponents);
        int i = nTeam[team]++;
        int s = order[i < 5 ? i : 4];
        CarS& car = a.cars[c];
        car_set_default(car, cfg.mut.carSpawnBoost);
        V3 pos(SX[s], SY[s], C::CAR_SPAWN_REST_Z);
        float yaw = SYAW[s];
        if (team == 1) { pos = V3(pos.x * -1.f, pos.y * -1.f, pos.z * 1.f); yaw = (float)((double)yaw + 3.14159265358979323846); }
        car_set_pose(car, pos, yaw, 0, 0);
    }
    a.ball.pos = V3(0.f * UU2BT, 0.f * UU2BT, C::BALL_REST_Z * UU2BT);
    a.ball.vel = V3(); a.ball.angvel = V3();
    a.ball.updateCounterLo = 0;
    pads_reset(a.pads);
}

RL_HDI V3 rand_vec(ArenaS& a, V3 lo, V3 hi) {
    float x = rng_float(a, lo.x, hi.x), y = rng_float(a, lo.y, hi.y), z = rng_float(a, lo.z, hi.z);
    return V3(x, y, z);
}

// RandomState::ResetState (G/Utils/StateSetters/RandomState.cpp:8-62)
RL_HDI void reset_to_random(ArenaS& a, const SimCfg& cfg) {
    reset_to_kickoff(a, cfg);
    const float X_MAX = 3500, Y_MAX = 4000, Z_MAX = 1820, CAR_Z_MIN = 150;
    const float PITCH_MAX = kPi / 2, YAW_MAX = kPi, ROLL_MAX = kPi, ANGVEL_MAX = 5.5f;
    {
        V3 p = rand_vec(a, V3(-X_MAX, -Y_MAX, 92.75f), V3(X_MAX, Y_MAX, Z_MAX));
        V3 v, w;
        if (cfg.randBallSpeed) {
            V3 d = ref_normalized(rand_vec(a, V3(-1, -1, -1), V3(1, 1, 1)));
            float s = rng_float(a, 0, 4000);
            v = V3(d.x * s, d.y * s, d.z * s);
            w = rand_vec(a, V3(-4, -4, -4), V3(4, 4, 4));
        }
        a.ball.pos = V3(p.x * UU2BT, p.y * UU2BT, p.z * UU2BT);
        a.ball.vel = V3(v.x * UU2BT, v.y * UU2BT, v.z * UU2BT);
        a.ball.angvel = w;
        a.ball.updateCounterLo = 0;
    }
    for (int ci = 0; ci < cfg.numCars; ci++) {
        CarS& car = a.cars[cfg.playerOrder[ci]];
        car_set_default(car, cfg.mut.carSpawnBoost);
        V3 p = rand_vec(a, V3(-X_MAX, -Y_MAX, CAR_Z_MIN), V3(X_MAX, Y_MAX, Z_MAX));
        V3 v, w;
        if (cfg.randCarSpeed) {
            (void)rand_vec(a, V3(-1, -1, -1), V3(1, 1, 1));  // unused randVelDir draw (RandomState.cpp:41)
            V3 d = ref_normalized(rand_vec(a, V3(-1, -1, -1), V3(1, 1, 1)));
            float s = rng_float(a, 0, C::CAR_MAX_SPEED);
            v = V3(d.x * s, d.y * s, d.z * s);
            V3 d2 = ref_normalized(rand_vec(a, V3(-1, -1, -1), V3(1, 1, 1)));
            w = V3(d2.x * ANGVEL_MAX, d2.y * ANGVEL_MAX, d2.z * ANGVEL_MAX);
        }
        float yaw = rng_float(a, -YAW_MAX, YAW_MAX), pitch = rng_float(a, -PITCH_MAX, PITCH_MAX), roll = rng_float(a, -ROLL_MAX, ROLL_MAX);
        bool onGround = cfg.carsOnGround ? true : (rng_float(a, 0, 1) > 0.5f);
        if (onGround) { p.z = 17; pitch = roll = 0; v.z = 0; w = V3(); }
        car_set_pose(car, p, yaw, pitch, roll);
        car.vel = V3(v.x * UU2BT, v.y * UU2BT, v.z * UU2BT);
        car.angvel = w;
        car.boost = rng_float(a, 0, 100);
    }
}

RL_HD RL_NOINLINE inline void gym_reset(ArenaS& a, const SimCfg& cfg) {
    if (cfg.stateSetter == 0) reset_to_kickoff(a, cfg);
    else if (cfg.stateSetter == 1) reset_to_random(a, cfg);
    episode_reset(a, cfg);
}

}  // namespace rl
