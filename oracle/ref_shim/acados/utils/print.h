/* TEST INFRASTRUCTURE ONLY - see acados_c/ocp_qp_interface.h in this directory. */
#ifndef DFKIREF_ACADOS_PRINT_H_
#define DFKIREF_ACADOS_PRINT_H_
#include "acados_c/ocp_qp_interface.h"
#ifdef __cplusplus
extern "C" {
#endif
void print_ocp_qp_dims(ocp_qp_dims *dims);
void print_qp_info(qp_info *info);
#ifdef __cplusplus
}
#endif
#endif
