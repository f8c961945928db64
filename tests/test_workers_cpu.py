"""The host worker pool of the staged host-memory path (csrc/t2fit_workers.h: spin-then-sleep) as a stand-alone C++
program: every task runs on every worker exactly once, back-to-back tasks and tasks after an idle period (sleeping
workers) both complete, destruction joins."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include "t2fit_workers.h"
#include <stdio.h>
#include <atomic>
#include <vector>
int main() {
    for (int nthreads : {1, 2, 5, 16}) {
        t2fit::Workers w(nthreads);
        std::vector<long long> acc(nthreads, 0);
        std::atomic<long long> calls{0};
        long long expect = 0;
        for (int it = 0; it < 6000; ++it) {
            if (it % 1500 == 1499) std::this_thread::sleep_for(std::chrono::milliseconds(3));   // workers go to sleep
            w.run([&](int part, int parts) {
                if (parts != nthreads) abort();
                acc[part] += it + part;
                calls.fetch_add(1);
            });
            for (int p = 0; p < nthreads; ++p) expect += it + p;
        }
        long long got = 0;
        for (long long v : acc) got += v;
        if (got != expect || calls.load() != 6000LL * nthreads) { printf("FAIL %d\n", nthreads); return 1; }
    }
    printf("OK\n");
    return 0;
}
'''


def test_worker_pool(tmp_path):
    src = tmp_path / "workers_test.cpp"
    src.write_text(SRC)
    exe = tmp_path / "workers_test"
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "fetal_t2mapping_b200", "csrc"), str(src), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True, timeout=120).stdout
    assert out.strip() == "OK"
