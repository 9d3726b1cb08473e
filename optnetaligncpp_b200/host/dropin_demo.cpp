// Small driver in the shape of the reference's hillclimber.cpp:52-79, written against the
// reference's UNCHANGED headers and linked against the drop-in Alignment.cpp / localSearch.cpp
// of this directory (plus the reference's own Network.cpp, blastinfo.cpp, goc.cpp, nsga-ii.cpp).
// It is what tests/test_gpu_dropin.py runs to show that reference callers work unchanged:
//
//   dropin_demo net1 net2 sim|- annos1|- annos2|- seed iters nthreads out_prefix [steepest]
//
// Per thread t: Alignment(net1,net2,bit,goc); shuf(mt19937(seed+t)); computeFitness;
// correctHillClimb(mt19937(seed+1000+t), iters, obj=-1); then mutate + a crossover with thread
// 0's result; prints counters/objectives and writes <out_prefix>_<t>.aln via Alignment::save.
// Last, one NSGA-II generation's ranking (nsgaiimain.cpp:157-175) of a small population of
// mutated children, once by the reference's own nsga-ii.cpp (linked unchanged) and once by
// ona_host::rankGeneration on copies; prints whether the two agree member by member.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <thread>
#include <vector>

#include "Alignment.h"
#include "blastinfo.h"
#include "goc.h"
#include "localSearch.h"
#include "nsga-ii.h"
#include "ona_host.hpp"

using namespace std;

int main(int argc, char** argv) {
    if (argc < 10) {
        cerr << "usage: dropin_demo net1 net2 sim|- annos1|- annos2|- seed iters nthreads out_prefix" << endl;
        return 2;
    }
    try {
        Network* net1 = new Network(argv[1]);
        Network* net2 = new Network(argv[2]);
        const string sim = argv[3], a1 = argv[4], a2 = argv[5];
        BLASTDict* bit = sim == "-" ? nullptr : loadBLASTInfo(net1, net2, sim);
        GOCDict* goc = (a1 == "-" || a2 == "-") ? nullptr : loadGOC(net1, net2, a1, a2);
        const int seed = atoi(argv[6]), iters = atoi(argv[7]), nthreads = atoi(argv[8]);
        const string prefix = argv[9];
        vector<fitnessName> names;  // argParsing.h:212-271 push order
        names.push_back(S3Fit);
        if (bit) names.push_back(BitscoreSumFit);
        if (goc) names.push_back(GOCFit);

        if (argc > 10 && string(argv[10]) == "smallops") {
            // throughput of the short per-alignment calls (mutate = one swap list, shuf = two
            // from-scratch evaluations) from `nthreads` threads, with and without a long batched
            // climb running beside them: the situation of steadystate.cpp's workers
            for (int with_climb = 0; with_climb < 2; with_climb++) {
                atomic<bool> stop(false);
                atomic<long long> nmut(0), nshuf(0);
                vector<thread> ts;
                for (int t = 0; t < nthreads; t++)
                    ts.push_back(thread([&, t]() {
                        RandGenT g(seed + t);
                        Alignment a(net1, net2, bit, goc);
                        a.shuf(g, false, false, true);
                        while (!stop) {
                            a.mutate(g, 0.05f, true);
                            nmut++;
                            a.shuf(g, false, false, true);
                            nshuf++;
                        }
                    }));
                thread climber;
                if (with_climb)
                    climber = thread([&]() {
                        vector<Alignment*> pop;
                        RandGenT g(seed + 999);
                        Alignment a(net1, net2, bit, goc);
                        a.shuf(g, false, false, true);
                        a.computeFitness(names);
                        while (!stop) correctHillClimb(g, &a, true, iters, names);
                    });
                this_thread::sleep_for(chrono::milliseconds(3000));
                stop = true;
                for (auto& th : ts) th.join();
                if (with_climb) climber.join();
                printf("smallops threads %d climb_beside %d: %.0f swap lists/s, %.0f evaluations/s\n", nthreads, with_climb,
                       nmut / 3.0, 2.0 * nshuf / 3.0);
            }
            return 0;
        }
        if (argc > 10 && string(argv[10]) == "steepest") {
            // the threaded exhaustive search (localSearch.cpp:122-203): every scan of the n1 x n2
            // grid is one pass of the swap-delta kernel here
            RandGenT g(seed);
            Alignment a(net1, net2, bit, goc);
            a.shuf(g, false, false, true);
            a.computeFitness(names);
            steepestAscentHillClimb(&a, names, nthreads, false);
            printf("steepest cons %d ind %d bits %.17g goc %.17g", a.currConservedCount, a.currInducedCount,
                   bit ? a.currBitscore : 0.0, goc ? a.currGOC : 0.0);
            for (double f : a.fitness) printf(" %.17g", f);
            printf("\n");
            a.save(prefix + "_steepest.aln");
            return 0;
        }
        vector<Alignment*> alns(nthreads, nullptr);
        vector<thread> ts;
        for (int t = 0; t < nthreads; t++) {
            ts.push_back(thread([&, t]() {
                RandGenT g(seed + t);
                Alignment* aln = new Alignment(net1, net2, bit, goc);
                aln->shuf(g, false, false, true);
                aln->computeFitness(names);
                RandGenT h(seed + 1000 + t);
                correctHillClimb(h, aln, true, iters, names);
                alns[t] = aln;
            }));
        }
        for (auto& th : ts) th.join();
        for (int t = 0; t < nthreads; t++) {
            Alignment* aln = alns[t];
            printf("climb %d cons %d ind %d bits %.17g goc %.17g", t, aln->currConservedCount,
                   aln->currInducedCount, bit ? aln->currBitscore : 0.0, goc ? aln->currGOC : 0.0);
            for (double f : aln->fitness) printf(" %.17g", f);
            printf("\n");
            aln->save(prefix + "_" + to_string(t) + ".aln");
        }
        ona_host::writeResults(prefix + "_res", alns, names);  // .info + .aln, steadystate.cpp:339-395
        // operators that precede the hot loop for every child (steadystate.cpp:222-232)
        for (int t = 0; t < nthreads; t++) {
            RandGenT g(seed + 2000 + t), m(seed + 3000 + t);
            Alignment child(g, 0.1f, *alns[t], *alns[0], true);
            child.mutate(m, 0.005f, true);
            child.computeFitness(names);
            printf("child %d cons %d ind %d bits %.17g goc %.17g ics %.17g", t, child.currConservedCount,
                   child.currInducedCount, bit ? child.currBitscore : 0.0, goc ? child.currGOC : 0.0, child.ics());
            for (double f : child.fitness) printf(" %.17g", f);
            printf("\n");
        }
        {
            vector<Alignment*> refPop, gpuPop;
            RandGenT g(seed + 4000);
            for (int i = 0; i < 48; i++) {
                Alignment* a = new Alignment(*alns[i % nthreads]);
                if (i >= nthreads) a->mutate(g, 0.002f * (1 + i % 5), true);
                a->computeFitness(names);
                refPop.push_back(a);
                gpuPop.push_back(new Alignment(*a));
            }
            bool same = true;
            size_t nfronts = 0;
            for (int gen = 0; gen < 2; gen++) {  // second pass: stale fitnessNormalized entries (:79)
                vector<vector<Alignment*> > fr = nonDominatedSort(refPop);
                normalizeFitnesses(refPop);
                for (auto& f : fr) setCrowdingDists(f);
                vector<vector<Alignment*> > fg = ona_host::rankGeneration(gpuPop);
                same = same && fr.size() == fg.size();
                nfronts = fg.size();
                for (size_t f = 0; same && f < fr.size(); f++) {
                    same = fr[f].size() == fg[f].size();
                    for (size_t k = 0; same && k < fr[f].size(); k++) {
                        const size_t ir = find(refPop.begin(), refPop.end(), fr[f][k]) - refPop.begin();
                        const size_t ig = find(gpuPop.begin(), gpuPop.end(), fg[f][k]) - gpuPop.begin();
                        same = ir == ig && fr[f][k]->domRank == fg[f][k]->domRank &&
                               fr[f][k]->crowdDist == fg[f][k]->crowdDist &&
                               fr[f][k]->fitnessNormalized == fg[f][k]->fitnessNormalized;
                    }
                }
                for (int i = 0; i < 10; i++) {  // next generation: replace a few members
                    refPop[i]->mutate(g, 0.004f, true);
                    refPop[i]->computeFitness(names);
                    refPop[i]->fitnessNormalized.clear();  // a new child has none yet
                    *gpuPop[i] = *refPop[i];
                }
            }
            printf("nsga2 fronts %zu %s\n", nfronts, same ? "identical" : "DIFFERENT");
        }
        {   // proportionalSearch (localSearch.cpp:60-84): 500-iteration climbs on one objective or another
            Alignment ps(*alns[0]);
            ps.computeFitness(names);
            RandGenT g(seed + 9000);
            proportionalSearch(g, &ps, true, 6, names, 0, 0.5);
            printf("propsearch cons %d ind %d bits %.17g goc %.17g", ps.currConservedCount, ps.currInducedCount,
                   bit ? ps.currBitscore : 0.0, goc ? ps.currGOC : 0.0);
            for (double f : ps.fitness) printf(" %.17g", f);
            printf("\n");
        }
        Alignment probe(*alns[0]);
        printf("snd %.17g\n", swapNormalizedDelta(probe, names, 3, 7));
        // partial alignments (no --total): random mask (shuf), mask flips (mutate), mask AND/OR
        // (UPMX), onBit, and a climb over the masked state
        auto report = [&](const char* tag, Alignment& a) {
            a.computeFitness(names);
            printf("%s cons %d ind %d bits %.17g goc %.17g size %.17g", tag, a.currConservedCount,
                   a.currInducedCount, bit ? a.currBitscore : 0.0, goc ? a.currGOC : 0.0, a.alnSize());
            for (double f : a.fitness) printf(" %.17g", f);
            printf("\n");
        };
        vector<Alignment*> parts;
        for (int t = 0; t < 2; t++) {
            RandGenT g(seed + 5000 + t), m(seed + 6000 + t);
            Alignment* a = new Alignment(net1, net2, bit, goc);
            a->shuf(g, false, false, false);
            report("pshuf", *a);
            a->mutate(m, 0.01f, false);
            report("pmut", *a);
            parts.push_back(a);
        }
        RandGenT c(seed + 7000), h(seed + 8000);
        Alignment child(c, 0.2f, *parts[0], *parts[1], false);
        report("pcx", child);
        int off = 0;
        while (off < child.actualSize && child.alnMask[off]) off++;
        child.onBit(off);
        printf("pbit %d ", off);
        report("pon", child);
        correctHillClimb(h, &child, false, iters, names);
        report("pclimb", child);
        // the other ways the drivers build a starting alignment (steadystate.cpp:92-118): greedy
        // best-score matching and the size-controlled random starts
        if (bit) {
            Alignment gm(net1, net2, bit, goc);
            gm.greedyMatch(true);
            report("greedy", gm);
            // --total --seeding: a greedyMatch individual (mask bits off on every position >= n1)
            // as a crossover parent; the reference leaves the mask bits in place under --total
            for (int v = 0; v < 4; v++) {
                RandGenT cg(seed + 9700 + v);
                Alignment gchild(cg, 0.3f, gm, *alns[0], true);
                report("gcx", gchild);
            }
        }
        for (int v = 0; v < 2; v++) {
            RandGenT g(seed + 9500 + v);
            Alignment a(net1, net2, bit, goc);
            a.shuf(g, v == 0, v == 1, false);
            report(v == 0 ? "uniformsize" : "smallstart", a);
        }
        // on-disk format either side of the path: save a partial alignment, load it back through
        // the alignment-file constructor (Alignment.cpp:49-138), which re-seeds the counters
        child.save(prefix + "_partial.aln");
        Alignment loaded(net1, net2, prefix + "_partial.aln", bit, goc);
        report("pload", loaded);
    } catch (exception& e) {
        cerr << "error: " << e.what() << endl;
        return 1;
    }
    return 0;
}
