// ber_bench.cpp -- command-line counterpart of the reference testbench (src/testbench/main.cpp):
// BPSK/AWGN/quantiser -> SC decoder -> error counter, all on the device, through the C ABI.
// Unlike main.cpp (no arguments, macros), everything is a run-time option.
//
//   ber_bench --order FILE | --flags FILE  -n N -k K [--par 16] [--q 8] [--sm] [--no-ext]
//             [--prune 0|1|2|3]  (3 = the reference's PRUNING_LEVEL 2 decoder: R0 / R1 / REP / SPC, not plain SC)
//             [--snr 2.5[:step:stop]] [--rate R] [--frames F] [--seed 0xF0] [--device D]
//             [--gpus G]   (frames split over G devices as independent streams; counters summed)
//             [--monitor]  (print the function x level matrix of sc_monitor.h for this table and exit; no GPU needed)
//             [--json]     (one JSON line per Eb/N0 point: counters, per-GPU counters, seconds, Gb/s)
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../include/scpd.hpp"

int main(int argc, char** argv) {
    std::string order, flagsf;
    uint32_t n = 0, k = 0, par = 16, q = 8, fmt = SCPD_FMT_CA2, ext = 1, prune = SCPD_PRUNE_R0_R1;
    double snr0 = 2.5, snr_step = 0.0, snr1 = 2.5, rate = -1.0;
    uint64_t frames = 1 << 16;
    int seed = 0xF0, device = 0, gpus = 1;
    bool monitor = false, json = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--order") order = next();
        else if (a == "--flags") flagsf = next();
        else if (a == "-n") n = (uint32_t)std::strtoul(next(), nullptr, 0);
        else if (a == "-k") k = (uint32_t)std::strtoul(next(), nullptr, 0);
        else if (a == "--par") par = (uint32_t)std::strtoul(next(), nullptr, 0);
        else if (a == "--q") q = (uint32_t)std::strtoul(next(), nullptr, 0);
        else if (a == "--sm") fmt = SCPD_FMT_SIGMAG;
        else if (a == "--no-ext") ext = 0;
        else if (a == "--prune") prune = (uint32_t)std::strtoul(next(), nullptr, 0);
        else if (a == "--rate") rate = std::atof(next());
        else if (a == "--frames") frames = std::strtoull(next(), nullptr, 0);
        else if (a == "--seed") seed = (int)std::strtol(next(), nullptr, 0);
        else if (a == "--device") device = std::atoi(next());
        else if (a == "--gpus") gpus = std::atoi(next());
        else if (a == "--monitor") monitor = true;
        else if (a == "--json") json = true;
        else if (a == "--snr") {
            double v[3] = {2.5, 0, 0};
            int c = std::sscanf(next(), "%lf:%lf:%lf", &v[0], &v[1], &v[2]);
            snr0 = v[0];
            snr_step = c == 3 ? v[1] : 0.0;
            snr1 = c == 3 ? v[2] : v[0];
        } else {
            std::fprintf(stderr, "unknown option %s\n", a.c_str());
            return 2;
        }
    }
    if (!n || (order.empty() && flagsf.empty())) {
        std::fprintf(stderr, "usage: ber_bench (--order FILE -k K | --flags FILE) -n N [options]\n");
        return 2;
    }
    try {
        std::vector<uint8_t> flags = order.empty() ? scpd::load_flag_table(flagsf, n, &k) : scpd::load_order_table(order, n, k);
        if (rate <= 0) rate = double(k) / double(n);  // the reference hard-codes R = 0.5 (main.cpp:92)
        scpd_config cfg{n, k, par, q, fmt, ext, prune, 0};
        if (monitor) {  // layout of sc_monitor.h:170-190 ("Matrix Level / Function"), loop iterations per frame
            scpd_stage_matrix m;
            if (scpd_stage_profile(&cfg, flags.data(), &m) != SCPD_OK) throw std::runtime_error("bad configuration");
            int top = 0;
            while ((1u << top) < n) top++;
            static const char* names[SCPD_STAGE_FUNCS] = {"F", "G", "H", "R", "R_R0", "R_R1", "F_REP", "G_SPC"};
            std::printf("[MONITOR] loop iterations per frame : %llu (pruning %u)\n", (unsigned long long)m.total_iterations, prune);
            std::printf("  Level     |");
            for (int l = top; l >= 0; l--) std::printf("%8u", 1u << l);
            std::printf("\n");
            for (int f = 0; f < SCPD_STAGE_FUNCS; f++) {
                std::printf("  fct %-5s |", names[f]);
                for (int l = top; l >= 0; l--)
                    std::printf("%8llu", (unsigned long long)(f >= SCPD_STAGE_R0 ? m.visits[f][l] : m.iterations[f][l]));
                std::printf("\n");
            }
            return 0;
        }
        if (!json) std::printf("(II) Frame size %u, K %u, LLR width %u, QUANT [4, -31, 31], PAR %u, %s, EXTENDED %u, %d GPU(s)\n", n, k,
                    q, par, fmt == SCPD_FMT_CA2 ? "CA2" : "SIGMAG", ext, gpus);
        // one handle and one host thread per GPU, created once; frames [lo, hi) of the stream per GPU; no collective
        std::vector<std::unique_ptr<scpd::PolarDecoder>> decs(gpus);
        for (int g = 0; g < gpus; g++) decs[g] = std::make_unique<scpd::PolarDecoder>(cfg, flags, device + g);
        auto run_point = [&](double snr, std::vector<scpd::BerCounters>& part) {
            std::vector<std::string> errs(gpus);
            std::vector<std::thread> th;
            for (int g = 0; g < gpus; g++) {
                th.emplace_back([&, g] {
                    try {
                        const uint64_t lo = frames * g / gpus, hi = frames * (g + 1) / gpus;
                        part[g] = decs[g]->run_ber((float)snr, (float)rate, hi - lo, lo, (uint8_t)seed);
                    } catch (const std::exception& e) {
                        errs[g] = e.what();
                    }
                });
            }
            for (auto& t : th) t.join();
            for (auto& e : errs)
                if (!e.empty()) throw std::runtime_error(e);
        };
        {   // untimed first pass: scratch buffers of the full batch, jump tables
            std::vector<scpd::BerCounters> warm(gpus);
            run_point(snr0, warm);
        }
        for (double snr = snr0; snr <= snr1 + 1e-9; snr += (snr_step > 0 ? snr_step : 1e9)) {
            std::vector<scpd::BerCounters> part(gpus);
            const auto t0 = std::chrono::steady_clock::now();
            run_point(snr, part);
            const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            scpd::BerCounters c{0, 0, 0, 0, 0, 0};
            for (auto& p : part) {  // the only "reduction" of the path: four (six) 64-bit counters per GPU
                c.bit_errors += p.bit_errors;
                c.frame_errors += p.frame_errors;
                c.bits += p.bits;
                c.frames += p.frames;
                c.bit_errors_wrapped += p.bit_errors_wrapped;
                c.frame_errors_wrapped += p.frame_errors_wrapped;
            }
            if (json) {
                std::printf("{\"n\": %u, \"k\": %u, \"gpus\": %d, \"ebn0_db\": %.2f, \"frames\": %llu, \"bit_errors\": %llu, "
                            "\"frame_errors\": %llu, \"bits\": %llu, \"seconds\": %.6f, \"info_gbps_incl_channel\": %.3f, \"per_gpu_bit_errors\": [",
                            n, k, gpus, snr, (unsigned long long)c.frames, (unsigned long long)c.bit_errors,
                            (unsigned long long)c.frame_errors, (unsigned long long)c.bits, sec, double(c.frames) * k / sec / 1e9);
                for (int g = 0; g < gpus; g++) std::printf("%s%llu", g ? ", " : "", (unsigned long long)part[g].bit_errors);
                std::printf("]}\n");
            } else {
                std::printf("Eb/N0 %.2f dB sigma %.4f | FRA %llu BE %llu FE %llu | BER %.3e FER %.3e | %.2f Gb/s info (incl. channel)\n",
                            snr, scpd_sigma((float)snr, (float)rate), (unsigned long long)c.frames,
                            (unsigned long long)c.bit_errors, (unsigned long long)c.frame_errors, c.ber(), c.fer(),
                            double(c.frames) * k / sec / 1e9);
            }
            std::fflush(stdout);
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
