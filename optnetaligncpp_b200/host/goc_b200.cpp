// Drop-in replacement for the reference's goc.cpp: implements the reference's UNCHANGED header goc.h.
// loadAnnotations parses exactly as the reference does (goc.cpp:55-89); loadGOC (goc.cpp:16-52) hands the
// pair loop -- O(n1 n2 terms), 31.8 s on one host thread for the bundled example pair -- to the GPU
// (ona_goc_matrix) and returns the same dense n1 x n2 GOCDict, value for value.
#include "goc.h"

#include <fstream>
#include <sstream>
#include <stdexcept>
#include <vector>

#include "ona_b200.h"

using namespace std;

// goc.cpp:55-89.  One line = a node name followed by integer GO terms.  As in the reference the name is
// looked up with operator[] on the network's name map, so a name the network does not know is
// default-inserted with id 0 and its set REPLACES node 0's (unknown names map to node 0, the
// last line wins); a line without terms gives an empty set, an empty line the name "".
unordered_map<node, set<int> > loadAnnotations(Network* net, string fp) {
    ifstream in(fp);
    if (!in.good()) throw LineReadException("Failed to load annotations! Check filename.");
    unordered_map<node, set<int> > result;
    string line;
    while (getline(in, line)) {
        istringstream fields(line);
        string name;
        fields >> name;
        set<int> terms;
        for (int term; fields >> term;) terms.insert(term);
        result[net->nodeNameToNode[name]] = terms;
    }
    return result;
}

namespace {
// the annotation map of one network as CSR over node ids (rows in std::set order = ascending, unique)
struct TermRows {
    vector<int> rowptr, terms;
    vector<unsigned char> present;
    TermRows(const unordered_map<node, set<int> >& anns, int n) : rowptr(n + 1, 0), present(n, 0) {
        for (const auto& kv : anns) {
            if (kv.first < 0 || kv.first >= n) throw runtime_error("annotation for a node outside the network");
            present[kv.first] = 1;
            rowptr[kv.first + 1] = (int)kv.second.size();
        }
        for (int u = 0; u < n; ++u) rowptr[u + 1] += rowptr[u];
        terms.resize(rowptr[n]);
        for (const auto& kv : anns) copy(kv.second.begin(), kv.second.end(), terms.begin() + rowptr[kv.first]);
    }
};
}  // namespace

// goc.cpp:16-52: Jaccard index card(A n B) / card(A u B) of the two term sets for every pair of annotated
// nodes (0/0 = NaN for two empty sets, :47), 0.0 where either node has no entry.
GOCDict* loadGOC(Network* net1, Network* net2, string file1, string file2) {
    const auto anns1 = loadAnnotations(net1, file1);
    const auto anns2 = loadAnnotations(net2, file2);
    const int n1 = (int)net1->nodeToNodeName.size(), n2 = (int)net2->nodeToNodeName.size();
    GOCDict* dict = new GOCDict(boost::extents[n1][n2]);
    if (n1 == 0 || n2 == 0) return dict;
    const TermRows r1(anns1, n1), r2(anns2, n2);
    int device = 0;
    if (const char* d = getenv("ONA_DEVICE")) device = atoi(d);
    const int rc = ona_goc_matrix(device, n1, n2, r1.rowptr.data(), r1.terms.data(), r1.present.data(),
                                  r2.rowptr.data(), r2.terms.data(), r2.present.data(), &(*dict)[0][0]);
    if (rc != ONA_OK) {
        const string msg = ona_last_error();
        delete dict;
        throw runtime_error("ona_b200: " + msg);
    }
    return dict;
}
