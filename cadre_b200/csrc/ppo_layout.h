// Flat fp32 layout of the 16 PPO modules (ppo_agent/models.py:44-126: 8 x Model + 8 x LSTM; 128 tensors,
// 19 382 808 parameters) shared by parameters, gradients and both Adam moments (expert e = head*4 + command, head 0 =
// steer, 1 = throttle). Rows of length 530 are padded to 532 floats (16-byte
// TMA stride rule); LSTM gate rows are interleaved (row 4*u+g holds gate g of hidden unit u, g = i,f,g,o) so
// that the LSTM-cell epilogue finds the four gates of a unit in adjacent accumulator columns. Padding elements
// are zero and stay zero (zero gradient => Adam leaves them untouched).
// cadre_b200/ppo_params.py mirrors these numbers and converts from / to reference state dicts.
#pragma once

namespace cadre {
namespace ppo {

constexpr int E = 8;        // experts
constexpr int F = 530;      // feature / hidden size (agent_config.py:20)
constexpr int LDF = 532;    // padded row length
constexpr int G = 4 * F;    // LSTM gate rows
constexpr int HID = 128;    // MLP hidden size (models.py:162)
constexpr int AMAX = 33;    // steer actions (throttle uses the first 3 rows)
constexpr int B3A_LD = 36;

// LSTM tensors are stored EXPERT-major: block e = [W_ih | W_hh | b_ih | b_hh] of expert e (one reference module,
// `<head>_lstm_<command>`), so that the gradient of a group of experts is ONE contiguous range of the flat buffer: a
// data-parallel learner all-reduces it and runs its clip + Adam while the weight-gradient GEMMs of the next group are
// still in flight (cadre_b200/learner.py). The actor-critic tensors follow, kind-major with equal expert strides (one
// grid.z-batched GEMM per layer covers all eight experts).
constexpr long long LSTM_WIH = 0;                              // offsets inside an expert's LSTM block
constexpr long long LSTM_WHH = (long long)G * LDF;
constexpr long long LSTM_BIH = 2LL * G * LDF;
constexpr long long LSTM_BHH = 2LL * G * LDF + G;
constexpr long long LSTM_BLK = 2LL * G * LDF + 2LL * G;        // 2 259 920 floats, a multiple of 4 (16-byte rows)
constexpr long long SZ_W1 = (long long)E * 2 * HID * LDF;   // rows 0..127 actor (control.linear.0), 128..255 critic.0
constexpr long long SZ_B1 = (long long)E * 2 * HID;
constexpr long long SZ_W2 = (long long)E * 2 * HID * HID;   // [e][0] control.linear.2, [e][1] critic.2
constexpr long long SZ_B2 = (long long)E * 2 * HID;
constexpr long long SZ_W3A = (long long)E * AMAX * HID;     // control.linear.4
constexpr long long SZ_B3A = (long long)E * B3A_LD;
constexpr long long SZ_W3C = (long long)E * HID;            // critic.4
constexpr long long SZ_B3C = (long long)E * 4;

constexpr long long OFF_LSTM = 0;
constexpr long long OFF_W1 = OFF_LSTM + E * LSTM_BLK;
constexpr long long OFF_B1 = OFF_W1 + SZ_W1;
constexpr long long OFF_W2 = OFF_B1 + SZ_B1;
constexpr long long OFF_B2 = OFF_W2 + SZ_W2;
constexpr long long OFF_W3A = OFF_B2 + SZ_B2;
constexpr long long OFF_B3A = OFF_W3A + SZ_W3A;
constexpr long long OFF_W3C = OFF_B3A + SZ_B3A;
constexpr long long OFF_B3C = OFF_W3C + SZ_W3C;
constexpr long long TOTAL = OFF_B3C + SZ_B3C;

}  // namespace ppo
}  // namespace cadre
