// Generates tests/golden/tch_varstore.ot: a VarStore-style checkpoint written the way tch's `Tensor::save_multi` does it
// (torch-sys `at_save_multi`: one torch::serialize::OutputArchive, `archive.write(name, tensor)` per named variable,
// `archive.save_to(path)`) -- the reference's NeRF::save (src/model.rs:211-213) goes through exactly that call.
// Ten nn::linear layers created on ONE path (src/model.rs:48-55, 89-90) collide on "bias"/"weight"; tch appends
// "__{number of variables so far}" to a colliding name, and nn::linear creates the bias before the weight.
// Tiny dims (hidden 6, in 3, heads like the as-shipped net); values are deterministic so the test can recompute them.
//   build: see tests/golden/make_ot_fixture.sh
#include <torch/torch.h>

#include <string>
#include <vector>

int main(int argc, char **argv) {
    const int W = 6, IN = 3, LABELS = 4;
    const int dims[10][2] = {{IN, W}, {W, W}, {W, W}, {W, W}, {W, W}, {W, W}, {W, W}, {W, W + 1}, {W, W / 2}, {W / 2, LABELS}};
    torch::serialize::OutputArchive archive;
    int n_vars = 0;
    float next = 0.f;
    for (int l = 0; l < 10; ++l) {
        const int in = dims[l][0], out = dims[l][1];
        auto name = [&](const char *base) { return n_vars < 2 ? std::string(base) : std::string(base) + "__" + std::to_string(n_vars); };
        torch::Tensor b = torch::arange(out, torch::kFloat32) * 0.5f + next;
        next += 100.f;
        archive.write(name("bias"), b.set_requires_grad(true));
        ++n_vars;
        torch::Tensor w = (torch::arange(out * in, torch::kFloat32) * 0.25f + next).reshape({out, in});
        next += 100.f;
        archive.write(name("weight"), w.set_requires_grad(true));
        ++n_vars;
    }
    archive.save_to(argc > 1 ? argv[1] : "tch_varstore.ot");
    return 0;
}
