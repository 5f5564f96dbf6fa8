#include <math.h>

#include <algorithm>

#include "ame_host.h"

namespace host {

// constants.h:94-103: lambda by effective QP.  Decimal literals are doubles converted to float, like the
// reference's `const float fullLambdas[60] = {...}` initialiser.
static const float kFullLambdas[60] = {
    0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0,
    2.769291, 3.108425, 3.489089, 3.916370, 4.395976, 4.934316, 5.538583, 6.216849, 6.978177,
    7.832739, 8.791952, 9.868633, 11.077166, 12.433698, 13.956355, 15.665478, 17.583905, 19.737266, 22.154332,
    24.867397, 27.912709, 31.330957, 35.167810, 39.474532, 44.308664, 49.734793, 55.825418, 62.661913, 70.335619,
    78.949063, 88.617327, 99.469587, 111.650836, 125.323826, 140.671239, 157.898127, 177.234655, 198.939174, 223.301672,
    250.647653, 281.342477, 315.796254, 354.469310, 397.878347, 446.603345, 501.295305, 562.684955, 631.592507, 708.938619};

int compute_delta_qp(int inputQp, int poc) {
    static const int pocOffset[8] = {1, 5, 4, 5, 4, 5, 4, 5};
    const bool key = (poc % 8) == 0;
    const double scale = key ? 0 : 0.259, offset = key ? 0 : -6.5;
    int qp = inputQp + pocOffset[poc % 8];
    const double d = qp * scale + offset + 0.5;
    qp += (int)floor(std::max(0.0, std::min(3.0, d)));
    return qp;
}

float lambda_for(int inputQp, int poc) {
    const int q = compute_delta_qp(inputQp, poc);
    return kFullLambdas[std::max(0, std::min(59, q))];
}

std::vector<std::vector<int>> reference_lists(int nFrames) {
    // Label-only replay of the 4-slot buffer rotation: newest at [0]; POC%8==0 frames become long-term
    // entries from the tail and are then only replaced by a newer POC%8==0 frame.
    int refs[4] = {-1, -1, -1, -1}, lt[4] = {0, 0, 0, 0};
    std::vector<std::vector<int>> out;
    for (int poc = 1; poc <= nFrames; poc++) {
        const int num = std::min(4, poc);
        int a = refs[0], b = -1;
        refs[0] = poc - 1;
        if (poc < 5) {
            if (num > 1) { b = refs[1]; refs[1] = a; }
            if (num > 2) { a = refs[2]; refs[2] = b; }
            if (num > 3) refs[3] = a;
            lt[3] = refs[3] % 8 == 0;
        } else {
            if (lt[1] == 0 || (a % 8 == 0 && a != refs[0])) {
                b = refs[1]; refs[1] = a;
                if (lt[2] == 0 || (b % 8 == 0 && b != refs[1])) {
                    a = refs[2]; refs[2] = b;
                    if (lt[3] == 0 || (a % 8 == 0 && a != refs[3])) refs[3] = a;
                }
            }
            lt[3] = refs[3] % 8 == 0;
            lt[2] = (refs[2] % 8 == 0) && lt[3];
            lt[1] = (refs[1] % 8 == 0) && lt[2];
        }
        out.emplace_back(refs, refs + num);
    }
    return out;
}

void print_reference_plan(int nFrames, int inputQp) {
    static const int pocOffset[8] = {1, 5, 4, 5, 4, 5, 4, 5};
    printf("-=-=-= Artificial references used for debugging =-=-=-=-\n");
    printf("Input QP = %d\n", inputQp);
    const auto lists = reference_lists(nFrames);
    for (int poc = 1; poc < nFrames; poc++) {  // the reference's loop stops one frame short (f < N_FRAMES)
        const int qp = compute_delta_qp(inputQp, poc);
        (void)pocOffset;
        printf("POC %3d   QP %d motionLambda %f : [L0", poc, qp, lambda_for(inputQp, poc));
        for (int r : lists[poc - 1]) printf(" %d", r);
        printf("]\n");
    }
}

}  // namespace host
