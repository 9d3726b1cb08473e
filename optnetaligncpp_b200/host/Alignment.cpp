// Drop-in replacement for the reference's Alignment.cpp: implements the reference's UNCHANGED
// header Alignment.h (found on the include path at build time, never copied) on top of the
// B200 engine's C ABI.  Every graph-touching computation -- from-scratch scoring, runs of
// doSwap, hill climbing -- happens on the GPU; what stays on the host is bookkeeping that
// never touches a network (permutation shuffles, PRNG draws, O(M) objective arithmetic, file
// I/O).  Each member cites the reference lines whose behaviour it reproduces
// (/root/reference/Alignment.cpp).
#include "Alignment.h"

#include <algorithm>
#include <cassert>
#include <fstream>
#include <sstream>
#include <unordered_set>

#include "ona_host.hpp"

using namespace std;

namespace {

void init_bookkeeping(Alignment& a) {
    a.fitnessValid = false;
    a.domRank = -1;
    a.numThatDominate = -1;
    a.crowdDist = -1.0;
}

// assign every still-unaligned position an arbitrary unaligned V2 node and mask it off
// (Alignment.cpp:116-127 / 321-330); iteration order of the unordered_set is the reference's
void fill_unaligned(Alignment& a, unordered_set<node>& v2Unaligned) {
    for (size_t i = 0; i < a.aln.size(); i++) {
        if (a.aln[i] != -1) continue;
        const node arb = *(v2Unaligned.begin());
        a.aln[i] = arb;
        a.alnInv[arb] = (node)i;
        a.alnMask[i] = false;
        v2Unaligned.erase(arb);
    }
}

}  // namespace

// Alignment.cpp:23-47 -- identity alignment, counters seeded from scratch
Alignment::Alignment(const Network* n1, const Network* n2, const BLASTDict* bit, const GOCDict* goc) {
    const size_t size = n2->nodeToNodeName.size();
    aln.resize(size);
    alnInv.resize(size);
    alnMask.assign(size, true);
    for (size_t i = 0; i < size; i++) aln[i] = alnInv[i] = (node)i;
    net1 = n1; net2 = n2; bitscores = bit; gocs = goc;
    actualSize = (int)n1->nodeToNodeName.size();
    init_bookkeeping(*this);
    currBitscore = 0.0;
    currGOC = 0.0;
    ona_host::seed_counters(*this);
}

// Alignment.cpp:49-138 -- load "name1 name2" lines; unknown names are skipped (:97-105)
Alignment::Alignment(const Network* n1, const Network* n2, string filename, const BLASTDict* bit,
                     const GOCDict* goc) {
    const int size = (int)n2->nodeToNodeName.size();
    aln.assign(size, -1);
    alnInv.assign(size, -1);
    alnMask.assign(size, true);
    net1 = n1; net2 = n2; bitscores = bit; gocs = goc;
    actualSize = (int)n1->nodeToNodeName.size();
    init_bookkeeping(*this);
    currBitscore = 0.0;
    currGOC = 0.0;
    ifstream infile(filename);
    if (!infile) throw LineReadException(string("Alignment file ") + filename + string(" failed to open!"));
    unordered_set<node> v2Unaligned;
    for (int i = 0; i < size; i++) v2Unaligned.insert(i);
    string line;
    while (getline(infile, line)) {
        istringstream iss(line);
        string a, b;
        if (!(iss >> a >> b))
            throw LineReadException(string("Parse error in network: ") + filename + string("on line: ") + line + "\n");
        if (!net1->nodeNameToNode.count(a) || !net2->nodeNameToNode.count(b)) continue;
        const node u = net1->nodeNameToNode.at(a), v = net2->nodeNameToNode.at(b);
        aln[u] = v;
        alnInv[v] = u;
        v2Unaligned.erase(v);
    }
    fill_unaligned(*this, v2Unaligned);
    ona_host::seed_counters(*this);
}

// Alignment.cpp:152-256 -- UPMX crossover.  The child starts as parent 1; for every index i
// chosen with probability cxswappb it takes parent 2's image and the displaced image moves to
// where that image was (:197-252).  With every mask bit on that is exactly
// doSwap(i, alnInv[par2.aln[i]]), so the ordered swap list is built on the host (it needs only
// the two permutations and the PRNG) and applied on the GPU with incremental counter updates;
// otherwise the step list is replayed by ona_apply_edits (ONA_EDIT_CROSS).
Alignment::Alignment(RandGenT& prng, float cxswappb, const Alignment& p1, const Alignment& p2, bool total) {
    uniform_int_distribution<int> coinFlip(0, 1);
    uniform_real_distribution<float> fltgen(0.0, 1.0);
    const Alignment* par1 = &p2;
    const Alignment* par2 = &p1;
    if (coinFlip(prng)) { par1 = &p1; par2 = &p2; }
    aln = par1->aln;
    alnInv = par1->alnInv;
    alnMask = par1->alnMask;
    init_bookkeeping(*this);
    bitscores = par1->bitscores; gocs = par1->gocs;
    net1 = par1->net1; net2 = par1->net2;
    actualSize = par1->actualSize;
    currBitscore = par1->currBitscore; currGOC = par1->currGOC;
    currConservedCount = par1->currConservedCount; currInducedCount = par1->currInducedCount;

    const int size = (int)aln.size();
    // Under --total the reference leaves the mask bits where they are (the mask writes sit behind
    // `if(!total)`, :219-227), whereas doSwap exchanges them (:502-504): the two only coincide when
    // every bit of parent 1 is on.  A greedyMatch individual (--total --seeding) or a loaded partial
    // alignment has bits off, so those parents take the edit path with the `total` flag set.
    const bool allOn = find(alnMask.begin(), alnMask.end(), false) == alnMask.end();
    if (total && !allOn) {
        vector<int> ops;
        for (int i = 0; i < size; i++) {
            if (!(fltgen(prng) < cxswappb)) continue;   // no coin flip under --total (:222)
            ops.push_back(ONA_EDIT_CROSS);
            ops.push_back(i);
            ops.push_back(par2->aln[i]);
            ops.push_back((par2->alnMask[i] ? 1 : 0) | 4);
        }
        ona_host::apply_edits(*this, ops);
    } else if (total) {
        vector<node> a = aln, inv = alnInv;   // scratch copies that track the running permutation
        vector<int> swaps;
        for (int i = 0; i < size; i++) {
            if (!(fltgen(prng) < cxswappb)) continue;
            const node want = par2->aln[i], j = inv[want], have = a[i];
            swaps.push_back(i);
            swaps.push_back(j);
            a[i] = want; inv[want] = i;
            a[j] = have; inv[have] = j;
        }
        ona_host::apply_swaps(*this, swaps);
    } else {
        // partial alignments: the mask of the swapped-in position becomes parent1 AND/OR parent2
        // (:219-227), which is not a pure doSwap.  The PRNG draws do not depend on the running
        // state, so the ordered step list is drawn here and replayed on the GPU (ONA_EDIT_CROSS).
        vector<int> ops;
        for (int i = 0; i < size; i++) {
            if (!(fltgen(prng) < cxswappb)) continue;
            const int coin = coinFlip(prng);
            ops.push_back(ONA_EDIT_CROSS);
            ops.push_back(i);
            ops.push_back(par2->aln[i]);
            ops.push_back((par2->alnMask[i] ? 1 : 0) | (coin ? 2 : 0));
        }
        ona_host::apply_edits(*this, ops);
    }
    fitnessValid = false;
}

// Alignment.cpp:260-342 -- greedy best-score matching; set iteration orders are libstdc++'s
void Alignment::greedyMatch(bool bit) {
    assert(!(bit && bitscores == nullptr));
    assert(!(!bit && gocs == nullptr));
    const BLASTDict* dict = bit ? bitscores : gocs;
    const int n1 = (int)net1->nodeToNodeName.size(), n2 = (int)net2->nodeToNodeName.size();
    aln.assign(n2, -1);
    alnInv = aln;
    unordered_set<node> v1Unaligned, v2Unaligned, v1Aligned;
    for (int i = 0; i < n1; i++) v1Unaligned.insert(i);
    for (int i = 0; i < n2; i++) v2Unaligned.insert(i);
    for (node u : v1Unaligned) {
        node best = -1;
        double bestScore = 0.0;
        for (node v : v2Unaligned) {
            if ((*dict)[u][v] > bestScore) { best = v; bestScore = (*dict)[u][v]; }
            else if (best == -1) best = v;
        }
        aln[u] = best;
        alnInv[best] = u;
        v1Aligned.insert(u);
        v2Unaligned.erase(best);
    }
    for (size_t i = 0; i < alnMask.size(); i++) alnMask[i] = v1Aligned.count((node)i) != 0;
    fill_unaligned(*this, v2Unaligned);
    ona_host::seed_counters(*this);
}

// Alignment.cpp:344-378 -- the reference's seedExtend only builds (and drops) a priority queue
void Alignment::seedExtend(bool bit, bool degdiff) {
    assert(!(bit && bitscores == nullptr));
    assert(!(!bit && gocs == nullptr));
    (void)degdiff;
}

// Alignment.cpp:380-447 -- random permutation (+ random mask unless --total), then from-scratch
// counters.  PRNG draws are made in the reference's order.
void Alignment::shuf(RandGenT& prng, bool uniformsize, bool smallstart, bool total) {
    shuffle(aln.begin(), aln.end(), prng);
    for (size_t i = 0; i < aln.size(); i++) alnInv[aln[i]] = (node)i;
    if (!total) {
        if (!uniformsize && !smallstart) {
            uniform_int_distribution<int> coinFlip(0, 1);
            for (size_t i = 0; i < alnMask.size(); i++)
                if (coinFlip(prng)) alnMask[i] = false;
        } else {
            int numToDeactivate;
            if (uniformsize) {
                uniform_int_distribution<int> sizeDist(1, actualSize - 1);
                numToDeactivate = sizeDist(prng);
            } else {
                int maxsize = (int)net1->nodeToNodeName.size() / 100;
                if (maxsize == 0) maxsize = 1;
                uniform_int_distribution<int> sizeDist(1, maxsize);
                numToDeactivate = actualSize - sizeDist(prng);
            }
            uniform_int_distribution<int> indexDist(0, actualSize - 1);
            for (int k = 0; k < numToDeactivate; k++) {
                int index = indexDist(prng);
                while (!alnMask[index]) index = indexDist(prng);
                alnMask[index] = false;
            }
            for (size_t i = actualSize; i < aln.size(); i++) alnMask[i] = false;
        }
    }
    ona_host::seed_counters(*this);
}

// Alignment.cpp:451-486 -- per index: maybe doSwap with a random other index, maybe (partial
// alignments only) flip the mask bit.  The swap list depends only on the PRNG, so under
// --total it is drawn on the host and applied on the GPU in one call.
void Alignment::mutate(RandGenT& prng, float mutswappb, bool total) {
    const int size = (int)aln.size();
    uniform_real_distribution<float> fltgen(0.0, 1.0);
    uniform_int_distribution<int> intgen(0, size - 2);
    if (total) {
        vector<int> swaps;
        for (int i = 0; i < size; i++) {
            if (fltgen(prng) < mutswappb) {
                int other = intgen(prng);
                if (other >= i) other++;
                swaps.push_back(i);
                swaps.push_back(other);
            }
        }
        ona_host::apply_swaps(*this, swaps);
    } else {
        vector<int> ops;
        for (int i = 0; i < size; i++) {
            if (fltgen(prng) < mutswappb) {
                int other = intgen(prng);
                if (other >= i) other++;
                ops.insert(ops.end(), {ONA_EDIT_SWAP, i, other, 0});
            }
            if (fltgen(prng) < mutswappb) ops.insert(ops.end(), {ONA_EDIT_FLIP, i, 0, 0});
        }
        ona_host::apply_edits(*this, ops);
    }
    fitnessValid = false;
}

// Alignment.cpp:491-520
void Alignment::doSwap(node x, node y) { ona_host::apply_swaps(*this, vector<int>{x, y}); }

// Alignment.cpp:522-533
void Alignment::onBit(node x) { ona_host::apply_edits(*this, vector<int>{ONA_EDIT_ONBIT, x, 0, 0}); }

// Alignment.cpp:535-556, 693-699 -- objectives from the cached counters
double Alignment::fastEC() const { return double(currConservedCount) / double(net1->edges.size()); }
double Alignment::fastS3() const {
    const double num = (double)currConservedCount;
    return num / (double(net1->edges.size() + currInducedCount) - num);
}
double Alignment::fastICS() const {
    const double denom = (double)currInducedCount;
    return denom == 0.0 ? 0.0 : double(currConservedCount) / denom;
}
double Alignment::icsTimesEC() const { return fastICS() * fastEC(); }
double Alignment::s3Variant() const {
    const double a = (double)currConservedCount, b = (double)(net1->edges.size()), c = (double)currInducedCount;
    return a / (max(b, c) - a);
}
double Alignment::fastICSNumerator() const { return (double)currConservedCount; }

// Alignment.cpp:559-602
void Alignment::computeFitness(const vector<fitnessName>& fitnessNames) {
    if (fitness.size() == 0) fitness = vector<double>(fitnessNames.size(), 0.0);
    for (size_t i = 0; i < fitnessNames.size(); i++) {
        double f = fitness[i];
        switch (fitnessNames[i]) {
            case ICSFit: f = fastICS(); break;
            case ECFit: f = fastEC(); break;
            case BitscoreSumFit: f = currBitscore; break;
            case EvalsSumFit: f = -1.0 * currBitscore; break;
            case SizeFit: f = alnSize(); break;
            case GOCFit: f = currGOC; break;
            case S3Fit: f = fastS3(); break;
            case S3DenomFit: f = -(double(net1->edges.size() + currInducedCount) - currConservedCount); break;
            case ICSTimesEC: f = icsTimesEC(); break;
            case S3Variant: f = s3Variant(); break;
        }
        fitness[i] = f;
    }
    fitnessValid = true;
}

// Alignment.cpp:604-661 -- set-based ICS: |conserved edge set| / |induced edge set|, counted on the GPU
double Alignment::ics() const { return ona_host::score_from_scratch(*this).ics; }
// Alignment.cpp:667-690
double Alignment::fastICSDenominator() const { return double(ona_host::score_from_scratch(*this).induced); }
// Alignment.cpp:701-722 (index-order fp64 sums)
double Alignment::sumBLAST() const { return ona_host::score_from_scratch(*this).bits; }
double Alignment::sumGOC() const { return ona_host::score_from_scratch(*this).goc; }
// Alignment.cpp:724-732
double Alignment::alnSize() const {
    double n = 0.0;
    for (int i = 0; i < actualSize; i++)
        if (alnMask[i]) n += 1.0;
    return n;
}

// Alignment.cpp:734-743
void Alignment::save(string filename) const {
    ofstream out(filename);
    for (size_t i = 0; i < net1->nodeToNodeName.size(); i++)
        if (alnMask[i]) out << net1->nodeToNodeName.at((node)i) << ' ' << net2->nodeToNodeName.at(aln[i]) << endl;
}

// Alignment.cpp:745-785 -- exported here (the reference declares them inline in the .cpp, so it
// emits no symbol at -O2; tests.cpp:115-116 only links at -O0)
void Alignment::updateBitscore(node n1, node n2old, node n2new, bool oldMask, bool newMask) {
    if (!bitscores || n1 >= actualSize) return;
    const double oldScore = oldMask ? (*bitscores)[n1][n2old] : 0.0;
    const double newScore = newMask ? (*bitscores)[n1][n2new] : 0.0;
    const double delta = newScore - oldScore;
    currBitscore += delta;
}
void Alignment::updateGOC(node n1, node n2old, node n2new, bool oldMask, bool newMask) {
    if (!gocs || n1 >= actualSize) return;
    const double oldScore = oldMask ? (*gocs)[n1][n2old] : 0.0;
    const double newScore = newMask ? (*gocs)[n1][n2new] : 0.0;
    currGOC += newScore - oldScore;
}

// Alignment.cpp:791-826 -- single-node queries (the reference's tests use them; the batched
// forms live in the engine's kernels)
int Alignment::conservedCount(node n1, node n2, bool mask, node ignore) const {
    if (!mask || n1 >= actualSize || n1 == ignore) return 0;
    int count = 0;
    for (node i : net1->adjList[n1])
        if (i != ignore && alnMask[i] && net2->adjMatrix[n2][aln[i]]) count++;
    return count;
}
int Alignment::inducedCount(node n2, node ignore) const {
    if (alnInv[n2] >= actualSize || !alnMask[alnInv[n2]]) return 0;
    int count = 0;
    for (node x : net2->adjList[n2])
        if (alnInv[x] < actualSize && alnMask[alnInv[x]] && x != ignore) count++;
    return count;
}
