// gemm_rows.cuh — internal names for the operand descriptors declared in include/gaddpg_b200.h.
#pragma once
#include "../../include/gaddpg_b200.h"

typedef gaddpg_operand Operand;
typedef gaddpg_nt_problem NTProblem;
typedef gaddpg_nt_group NTGroup;
typedef gaddpg_tn_problem TNProblem;
enum { OP_PLAIN = GADDPG_OP_PLAIN, OP_BNRELU = GADDPG_OP_BNRELU, OP_BNBWD = GADDPG_OP_BNBWD, OP_BNBWD_POOL = GADDPG_OP_BNBWD_POOL };
enum { EPI_STORE = GADDPG_EPI_STORE, EPI_DMASK = GADDPG_EPI_DMASK };
