// Test tool: the reference's checkpoint I/O for one network, with libtorch itself —
// torch::load / torch::save of a torch::nn::Sequential of Linear + ReLU layers, exactly what TorchLoadSaveSeq does
// (RLGymPPO_CPP/src/private/RLGymPPO_CPP/PPO/PPOLearner.cpp:372-419) on the model DiscretePolicy / ValueEstimator build
// (PPO/DiscretePolicy.cpp:11-27).
//   ckpt_roundtrip load <file.lt> <in> <h1,h2,..> <out>   -> prints sum / sum of squares of every parameter, in order
//   ckpt_roundtrip save <file.lt> <in> <h1,h2,..> <out>   -> writes a model whose parameter i is filled with (i + 1) / 8
#include <torch/torch.h>

#include <fstream>
#include <iostream>
#include <sstream>

static torch::nn::Sequential make(int in, const std::vector<int>& hidden, int out) {
    torch::nn::Sequential seq;
    int prev = in;
    for (int h : hidden) {
        seq->push_back(torch::nn::Linear(prev, h));
        seq->push_back(torch::nn::ReLU());
        prev = h;
    }
    seq->push_back(torch::nn::Linear(prev, out));
    return seq;
}

int main(int argc, char** argv) {
    if (argc != 6) return 2;
    std::string mode = argv[1], path = argv[2];
    int in = std::stoi(argv[3]), out = std::stoi(argv[5]);
    std::vector<int> hidden;
    std::stringstream ss(argv[4]);
    for (std::string tok; std::getline(ss, tok, ',');) hidden.push_back(std::stoi(tok));
    auto seq = make(in, hidden, out);
    torch::NoGradGuard ng;
    if (mode == "load") {
        std::ifstream f(path, std::ios::binary);
        f >> std::noskipws;
        if (!f.good()) return 3;
        torch::load(seq, f, torch::kCPU);
        std::cout.precision(9);
        for (auto& p : seq->parameters()) std::cout << p.numel() << " " << p.sum().item<double>() << " " << (p * p).sum().item<double>() << "\n";
    } else {
        int i = 0;
        for (auto& p : seq->parameters()) p.fill_((float)(++i) / 8.f);
        std::ofstream f(path, std::ios::binary);
        torch::save(seq, f);
    }
    return 0;
}
