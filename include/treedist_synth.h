/*
 * treedist_synth.h -- synthetic tree sets for bench.py and the parity tests; tooling, not part of the drop-in
 * boundary.  Exported by libtreedist_cuda.so so that the 100,000 x 200-taxon set of BASELINE config 4 is generated in under a
 * second; treecl_b200/synth.py implements the same two models in NumPy for small sets.
 *
 *   kind 0 (uniform)     emulates treeCl.tree.RandomTree.new(n, rooted=False) + randomise_branch_lengths()
 *                        (treeCl/tree.py:1234-1276, 1003-1025)
 *   kind 1 (structured)  emulates Simulator.generate_class_trees (treeCl/simulator.py:127-190): n_classes class trees, each the
 *                        master tree + class_nni random NNIs; gene tree k = class k % n_classes + Poisson(lam) NNIs
 * Taxon labels are t000 ... (zero padded to >= 3 digits), as treecl_b200.synth.taxon_labels.
 */
#ifndef TREEDIST_SYNTH_H
#define TREEDIST_SYNTH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Writes n_trees NUL-terminated Newick strings back to back into a malloc'ed buffer (*out_buffer, *out_bytes) - the layout
 * td_encode_newick_joined takes.  Tree k depends only on (seed, k, parameters), not on n_trees or n_threads.  Free with td_synth_free. */
int td_synth_trees(int kind, int64_t n_trees, int n_taxa, uint64_t seed, int n_classes, int class_nni, double lam, int n_threads,
                   char **out_buffer, int64_t *out_bytes);
void td_synth_free(char *buffer);

#ifdef __cplusplus
}
#endif
#endif /* TREEDIST_SYNTH_H */
