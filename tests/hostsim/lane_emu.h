// Lane emulator for the cooperative (lane-group-per-voxel) solver -- TEST TOOL ONLY.
//
// csrc/t2fit_lbfgsb_coop.cuh is SPMD code: G lanes run the same function, meet at group barriers (__syncwarp) and hand
// values round with shuffles.  Here every lane is a fiber (ucontext); a barrier switches to the next lane of a fixed
// ring, so between two barriers the lanes run one after the other -- forwards (lane 0 first) or backwards (lane G-1
// first).  Code that is free of races between barriers gives the same result in both orders; a lane that reads a value
// another lane writes in the same barrier interval (a missing barrier) sees the old value in one order and the new one in
// the other, which the tests turn into a bitwise mismatch.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <ucontext.h>

#include <functional>
#include <stdexcept>
#include <vector>

namespace t2fit {
namespace emu {

class Lanes {
public:
    Lanes(int g, bool reverse) : g_(g), ctx_(g), stacks_(g), barriers_(g, 0), dbox_(g, 0.0), bbox_(g, 0) {
        for (int p = 0; p < g; ++p) order_.push_back(reverse ? g - 1 - p : p);
        pos_of_.resize(g);
        for (int p = 0; p < g; ++p) pos_of_[order_[p]] = p;
        for (auto& s : stacks_) s = (char*)malloc(kStack);
    }
    ~Lanes() { for (auto s : stacks_) free(s); }

    // run body(lane) on all lanes; returns when every lane has returned
    void run(const std::function<void(int)>& body) {
        body_ = &body;
        for (int l = 0; l < g_; ++l) {
            barriers_[l] = 0;
            getcontext(&ctx_[l]);
            ctx_[l].uc_stack.ss_sp = stacks_[l];
            ctx_[l].uc_stack.ss_size = kStack;
            ctx_[l].uc_link = nullptr;
            const uintptr_t self = (uintptr_t)this;
            makecontext(&ctx_[l], (void (*)())trampoline, 3, (unsigned)(self & 0xffffffffu), (unsigned)(self >> 32), l);
        }
        swapcontext(&main_, &ctx_[order_[0]]);
        for (int l = 1; l < g_; ++l)
            if (barriers_[l] != barriers_[0]) throw std::runtime_error("lane emulator: lanes passed different numbers of barriers");
    }

    void barrier(int lane) {
        ++barriers_[lane];
        const int p = pos_of_[lane];
        swapcontext(&ctx_[lane], &ctx_[order_[(p + 1) % g_]]);
    }
    double shfl(int lane, double v, int src) {
        dbox_[lane] = v;
        barrier(lane);
        const double r = dbox_[src];
        barrier(lane);
        return r;
    }
    bool any(int lane, bool p) {
        bbox_[lane] = p ? 1 : 0;
        barrier(lane);
        int a = 0;
        for (int l = 0; l < g_; ++l) a |= bbox_[l];
        barrier(lane);
        return a != 0;
    }
    int width() const { return g_; }

private:
    static constexpr size_t kStack = 1 << 20;
    static void trampoline(unsigned lo, unsigned hi, int lane) {
        Lanes* self = (Lanes*)(((uintptr_t)hi << 32) | (uintptr_t)lo);
        (*self->body_)(lane);
        // a lane that has returned hands over to the next one of the ring; the last one returns to run()
        const int p = self->pos_of_[lane];
        if (p + 1 < self->g_) setcontext(&self->ctx_[self->order_[p + 1]]);
        setcontext(&self->main_);
    }
    int g_;
    std::vector<ucontext_t> ctx_;
    std::vector<char*> stacks_;
    std::vector<long> barriers_;
    std::vector<double> dbox_;
    std::vector<int> bbox_;
    std::vector<int> order_, pos_of_;
    ucontext_t main_;
    const std::function<void(int)>* body_ = nullptr;
};

}  // namespace emu
}  // namespace t2fit
