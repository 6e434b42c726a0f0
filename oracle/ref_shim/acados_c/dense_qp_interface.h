/* TEST INFRASTRUCTURE ONLY - see acados_c/ocp_qp_interface.h in this directory. */
#include "acados_c/ocp_qp_interface.h"
