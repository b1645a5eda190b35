// rl_convert.h — host AoS structs of the C ABI (include/rlgym_b200.h) <-> device ArenaS.
// Mirrors Car::SetState / Car::GetState / Ball::SetState / Ball::GetState
// (R/Sim/Car/Car.cpp:10-36, R/Sim/Ball/Ball.cpp:27-49): uu <-> Bullet units by * (1/50) and * 50.
#pragma once
#include "../../include/rlgym_b200.h"
#include "rl_state.h"

namespace rl {

RL_HDI V3 v3_from(const float* p) { return V3(p[0], p[1], p[2]); }
RL_HDI void v3_to(float* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
RL_HDI Controls controls_from(const rlg_controls& c) { return Controls{c.throttle, c.steer, c.pitch, c.yaw, c.roll, c.jump != 0, c.boost != 0, c.handbrake != 0}; }
RL_HDI void controls_to(rlg_controls& o, const Controls& c) {
    o.throttle = c.throttle; o.steer = c.steer; o.pitch = c.pitch; o.yaw = c.yaw; o.roll = c.roll;
    o.jump = c.jump; o.boost = c.boost; o.handbrake = c.handbrake;
}

RL_HD inline void car_from_pod(CarS& c, const rlg_car_state& i) {
    V3 pos = v3_from(i.pos), vel = v3_from(i.vel);
    c.pos = V3(pos.x * UU2BT, pos.y * UU2BT, pos.z * UU2BT);
    c.vel = V3(vel.x * UU2BT, vel.y * UU2BT, vel.z * UU2BT);
    c.angvel = v3_from(i.ang_vel);
    c.rot = M3::from_cols(v3_from(i.rot_forward), v3_from(i.rot_right), v3_from(i.rot_up));
    c.isOnGround = i.is_on_ground != 0;
    for (int k = 0; k < 4; k++) c.wheelContact[k] = i.wheels_with_contact[k] != 0;
    c.hasJumped = i.has_jumped != 0; c.hasDoubleJumped = i.has_double_jumped != 0; c.hasFlipped = i.has_flipped != 0;
    c.flipRelTorque = v3_from(i.flip_rel_torque);
    c.jumpTime = i.jump_time; c.flipTime = i.flip_time;
    c.isFlipping = i.is_flipping != 0; c.isJumping = i.is_jumping != 0;
    c.airTime = i.air_time; c.airTimeSinceJump = i.air_time_since_jump;
    c.boost = i.boost; c.timeSpentBoosting = i.time_spent_boosting;
    c.isSupersonic = i.is_supersonic != 0; c.supersonicTime = i.supersonic_time; c.handbrakeVal = i.handbrake_val;
    c.isAutoFlipping = i.is_auto_flipping != 0; c.autoFlipTimer = i.auto_flip_timer; c.autoFlipTorqueScale = i.auto_flip_torque_scale;
    c.worldContactHas = i.world_contact_has != 0; c.worldContactNormal = v3_from(i.world_contact_normal);
    c.carContactOtherId = i.car_contact_other_id; c.carContactCooldown = i.car_contact_cooldown;
    c.isDemoed = i.is_demoed != 0; c.demoRespawnTimer = i.demo_respawn_timer;
    c.hitValid = i.hit_valid != 0;
    c.hitRelPos = v3_from(i.hit_rel_pos_on_ball); c.hitBallPos = v3_from(i.hit_ball_pos); c.hitExtraVel = v3_from(i.hit_extra_vel);
    set_i64(c.hitTickLo, c.hitTickHi, i.hit_tick);
    set_i64(c.hitExtraTickLo, c.hitExtraTickHi, i.hit_extra_tick);
    c.lastControls = controls_from(i.last_controls);
    c.wheelSteer = i.wheel_steer_angle; c.wheelEngine = i.wheel_engine_force; c.wheelBrake = i.wheel_brake;
    for (int k = 0; k < 4; k++) { c.wheelLat[k] = i.wheel_lat_friction[k]; c.wheelLong[k] = i.wheel_long_friction[k]; c.wheelPush[k] = i.wheel_extra_pushback[k]; }
}

RL_HD inline void car_to_pod(rlg_car_state& o, const CarS& c, int carIndex, int spawnOpponents) {
    v3_to(o.pos, V3(c.pos.x * BT2UU, c.pos.y * BT2UU, c.pos.z * BT2UU));
    v3_to(o.vel, V3(c.vel.x * BT2UU, c.vel.y * BT2UU, c.vel.z * BT2UU));
    v3_to(o.ang_vel, c.angvel);
    v3_to(o.rot_forward, c.rot.col(0)); v3_to(o.rot_right, c.rot.col(1)); v3_to(o.rot_up, c.rot.col(2));
    o.is_on_ground = c.isOnGround;
    for (int k = 0; k < 4; k++) o.wheels_with_contact[k] = c.wheelContact[k];
    o.has_jumped = c.hasJumped; o.has_double_jumped = c.hasDoubleJumped; o.has_flipped = c.hasFlipped;
    v3_to(o.flip_rel_torque, c.flipRelTorque);
    o.jump_time = c.jumpTime; o.flip_time = c.flipTime; o.is_flipping = c.isFlipping; o.is_jumping = c.isJumping;
    o.air_time = c.airTime; o.air_time_since_jump = c.airTimeSinceJump;
    o.boost = c.boost; o.time_spent_boosting = c.timeSpentBoosting;
    o.is_supersonic = c.isSupersonic; o.supersonic_time = c.supersonicTime; o.handbrake_val = c.handbrakeVal;
    o.is_auto_flipping = c.isAutoFlipping; o.auto_flip_timer = c.autoFlipTimer; o.auto_flip_torque_scale = c.autoFlipTorqueScale;
    o.world_contact_has = c.worldContactHas; v3_to(o.world_contact_normal, c.worldContactNormal);
    o.car_contact_other_id = c.carContactOtherId; o.car_contact_cooldown = c.carContactCooldown;
    o.is_demoed = c.isDemoed; o.demo_respawn_timer = c.demoRespawnTimer;
    o.hit_valid = c.hitValid;
    v3_to(o.hit_rel_pos_on_ball, c.hitRelPos); v3_to(o.hit_ball_pos, c.hitBallPos); v3_to(o.hit_extra_vel, c.hitExtraVel);
    o.hit_tick = get_i64(c.hitTickLo, c.hitTickHi);
    o.hit_extra_tick = get_i64(c.hitExtraTickLo, c.hitExtraTickHi);
    controls_to(o.last_controls, c.lastControls);
    o.wheel_steer_angle = c.wheelSteer; o.wheel_engine_force = c.wheelEngine; o.wheel_brake = c.wheelBrake;
    for (int k = 0; k < 4; k++) { o.wheel_lat_friction[k] = c.wheelLat[k]; o.wheel_long_friction[k] = c.wheelLong[k]; o.wheel_extra_pushback[k] = c.wheelPush[k]; }
    o.car_id = carIndex + 1;
    o.team = car_team(carIndex, spawnOpponents);
}

RL_HDI void ball_from_pod(BallS& b, const rlg_ball_state& i) {
    b.pos = V3(i.pos[0] * UU2BT, i.pos[1] * UU2BT, i.pos[2] * UU2BT);
    b.vel = V3(i.vel[0] * UU2BT, i.vel[1] * UU2BT, i.vel[2] * UU2BT);
    b.angvel = v3_from(i.ang_vel);
    b.updateCounterLo = 0;  // Ball::SetState
}
RL_HDI void ball_to_pod(rlg_ball_state& o, const BallS& b) {
    v3_to(o.pos, V3(b.pos.x * BT2UU, b.pos.y * BT2UU, b.pos.z * BT2UU));
    v3_to(o.vel, V3(b.vel.x * BT2UU, b.vel.y * BT2UU, b.vel.z * BT2UU));
    v3_to(o.ang_vel, b.angvel);
}

RL_HDI void pad_from_pod(PadsS& p, int i, const rlg_pad_state& s) { pad_set(p, i, s.is_active != 0, s.cooldown, s.prev_locked_car_id); }
RL_HDI void pad_to_pod(rlg_pad_state& o, const PadsS& p, int i) {
    o.is_active = (int32_t)((pads_active(p) >> i) & 1ULL); o.cooldown = p.cooldown[i]; o.prev_locked_car_id = pad_locked(p, i);
}

// fresh arena: what Arena::Create + AddCar leaves behind (cars respawned, wheel carry-over zero)
RL_HD inline void arena_init(ArenaS& a, int numCars, uint64_t seed, uint64_t globalArenaId, float spawnBoost) {
    uint32_t* w = (uint32_t*)&a;
    for (size_t i = 0; i < sizeof(ArenaS) / 4; i++) w[i] = 0;
    uint64_t s = seed * 0x9E3779B97F4A7C15ULL + globalArenaId * 0xD1B54A32D192ED03ULL + 0x2545F4914F6CDD1DULL;
    a.rngLo = (uint32_t)s; a.rngHi = (uint32_t)(s >> 32);
    (void)rng_next(a); (void)rng_next(a);
    a.ball.pos = V3(0, 0, C::BALL_REST_Z * UU2BT);
    pads_reset(a.pads);
    a.lastTouchCarId = -1;
    for (int c = 0; c < numCars; c++) {
        CarS& car = a.cars[c];
        car.rot = M3::identity();
        car.pos = V3(0, 0, C::CAR_SPAWN_REST_Z * UU2BT);
        car.isOnGround = 1;
        car.boost = spawnBoost;  // Gym.cpp:42-48: SetMutatorConfig runs before AddCar -> Respawn(carSpawnBoostAmount)
        car.hitTickLo = car.hitTickHi = car.hitExtraTickLo = car.hitExtraTickHi = -1;
    }
}

}  // namespace rl
