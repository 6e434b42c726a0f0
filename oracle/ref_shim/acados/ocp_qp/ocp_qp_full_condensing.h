/* TEST INFRASTRUCTURE ONLY - see acados_c/ocp_qp_interface.h in this directory. */
#ifndef DFKIREF_OCP_QP_FULL_CONDENSING_H_
#define DFKIREF_OCP_QP_FULL_CONDENSING_H_
#include "acados_c/ocp_qp_interface.h"
typedef struct ocp_qp_full_condensing_memory { dense_qp_in *fcond_qp_in; } ocp_qp_full_condensing_memory;
#endif
