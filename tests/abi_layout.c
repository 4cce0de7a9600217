/* Prints the layout of every structure of include/massiv_b200.h as "struct.field offset size" lines; the test
 * tests/test_abi_layout.py compares them with the ctypes mirror (massiv_b200/_abi.py), the oracle's mirror and the
 * table in INTEGRATION.md section 2 — the offsets a Haskell binding pokes at. */
#include <stddef.h>
#include <stdio.h>

#include "massiv_b200.h"

#define F(S, f) printf(#S "." #f " %zu %zu\n", offsetof(S, f), sizeof(((S *)0)->f))

int main(void) {
  printf("mb_stencil_desc.sizeof %zu 0\n", sizeof(mb_stencil_desc));
  F(mb_stencil_desc, kind); F(mb_stencil_desc, elem); F(mb_stencil_desc, rank); F(mb_stencil_desc, border);
  F(mb_stencil_desc, size); F(mb_stencil_desc, center); F(mb_stencil_desc, pad_lo); F(mb_stencil_desc, pad_hi);
  F(mb_stencil_desc, stride); F(mb_stencil_desc, fill); F(mb_stencil_desc, coeffs); F(mb_stencil_desc, coeffs2);
  F(mb_stencil_desc, ntaps); F(mb_stencil_desc, tap_offsets); F(mb_stencil_desc, flags); F(mb_stencil_desc, n_post);
  F(mb_stencil_desc, post); F(mb_stencil_desc, n_terms); F(mb_stencil_desc, n_program); F(mb_stencil_desc, terms);
  F(mb_stencil_desc, program);
  printf("mb_term.sizeof %zu 0\n", sizeof(mb_term));
  F(mb_term, kind); F(mb_term, size); F(mb_term, center); F(mb_term, coeffs); F(mb_term, ntaps); F(mb_term, tap_offsets);
  printf("mb_xins.sizeof %zu 0\n", sizeof(mb_xins));
  F(mb_xins, op); F(mb_xins, arg); F(mb_xins, operand);
  printf("mb_post_op.sizeof %zu 0\n", sizeof(mb_post_op));
  F(mb_post_op, op); F(mb_post_op, operand);
  printf("mb_shard_plan.sizeof %zu 0\n", sizeof(mb_shard_plan));
  F(mb_shard_plan, first_plane); F(mb_shard_plan, local_planes); F(mb_shard_plan, halo_lo); F(mb_shard_plan, halo_hi);
  F(mb_shard_plan, recv_lower_from); F(mb_shard_plan, recv_upper_from); F(mb_shard_plan, send_first_to);
  F(mb_shard_plan, send_last_to); F(mb_shard_plan, ghost_lo_src); F(mb_shard_plan, ghost_hi_src);
  F(mb_shard_plan, step_halo_lo); F(mb_shard_plan, step_halo_hi);
  return 0;
}
