"""ctypes mirror of OSQP's public C API structs.

Layouts follow /root/reference/include/public/osqp_api_types.h:37-187 for a
non-packed, 32-bit-int build (OSQPInt = int).  OSQPFloat is double in the
parity/default build and float in the `f32` build.  The same mirror is used for
the B200 library (the product) and, in tests only, for the CPU oracle library,
because both link the unchanged reference core and so export the same public API.
"""
import ctypes as C

c_int = C.c_int

# enum osqp_linsys_solver_type  (osqp_api_constants.h:56-60)
OSQP_UNKNOWN_SOLVER = 0
OSQP_DIRECT_SOLVER = 1
OSQP_INDIRECT_SOLVER = 2
# osqp_precond_type (osqp_api_constants.h:66-69)
OSQP_NO_PRECONDITIONER = 0
OSQP_DIAGONAL_PRECONDITIONER = 1
# enum osqp_status_type (osqp_api_constants.h:28-40)
OSQP_SOLVED = 1
OSQP_SOLVED_INACCURATE = 2
OSQP_PRIMAL_INFEASIBLE = 3
OSQP_PRIMAL_INFEASIBLE_INACCURATE = 4
OSQP_DUAL_INFEASIBLE = 5
OSQP_DUAL_INFEASIBLE_INACCURATE = 6
OSQP_MAX_ITER_REACHED = 7
OSQP_TIME_LIMIT_REACHED = 8
OSQP_NON_CVX = 9
OSQP_SIGINT = 10
OSQP_UNSOLVED = 11
OSQP_INFTY = 1e30  # osqp_api_constants.h:196-203 (non CUDA+float value)
OSQP_NAN = float(0x7fc00000)  # osqp_api_constants.h:192-194: a finite marker value, not an IEEE NaN


def make_types(c_float):
    """Build the struct classes for a given OSQPFloat ctype."""

    class OSQPCscMatrix(C.Structure):
        _fields_ = [
            ("m", c_int), ("n", c_int),
            ("p", C.POINTER(c_int)), ("i", C.POINTER(c_int)),
            ("x", C.POINTER(c_float)),
            ("nzmax", c_int), ("nz", c_int), ("owned", c_int),
        ]

    class OSQPSettings(C.Structure):
        _fields_ = [
            ("device", c_int), ("linsys_solver", c_int),
            ("allocate_solution", c_int), ("verbose", c_int),
            ("profiler_level", c_int), ("warm_starting", c_int),
            ("scaling", c_int), ("polishing", c_int),
            ("rho", c_float), ("rho_is_vec", c_int),
            ("sigma", c_float), ("alpha", c_float),
            ("cg_max_iter", c_int), ("cg_tol_reduction", c_int),
            ("cg_tol_fraction", c_float), ("cg_precond", c_int),
            ("adaptive_rho", c_int), ("adaptive_rho_interval", c_int),
            ("adaptive_rho_fraction", c_float),
            ("adaptive_rho_tolerance", c_float),
            ("max_iter", c_int),
            ("eps_abs", c_float), ("eps_rel", c_float),
            ("eps_prim_inf", c_float), ("eps_dual_inf", c_float),
            ("scaled_termination", c_int), ("check_termination", c_int),
            ("check_dualgap", c_int), ("time_limit", c_float),
            ("delta", c_float), ("polish_refine_iter", c_int),
        ]

    class OSQPInfo(C.Structure):
        _fields_ = [
            ("status", C.c_char * 32), ("status_val", c_int),
            ("status_polish", c_int),
            ("obj_val", c_float), ("dual_obj_val", c_float),
            ("prim_res", c_float), ("dual_res", c_float),
            ("duality_gap", c_float),
            ("iter", c_int), ("rho_updates", c_int),
            ("rho_estimate", c_float),
            ("setup_time", c_float), ("solve_time", c_float),
            ("update_time", c_float), ("polish_time", c_float),
            ("run_time", c_float),
            ("primdual_int", c_float), ("rel_kkt_error", c_float),
        ]

    class OSQPSolution(C.Structure):
        _fields_ = [
            ("x", C.POINTER(c_float)), ("y", C.POINTER(c_float)),
            ("prim_inf_cert", C.POINTER(c_float)),
            ("dual_inf_cert", C.POINTER(c_float)),
        ]

    class OSQPSolver(C.Structure):
        _fields_ = [
            ("settings", C.POINTER(OSQPSettings)),
            ("solution", C.POINTER(OSQPSolution)),
            ("info", C.POINTER(OSQPInfo)),
            ("work", C.c_void_p),
        ]

    class T:
        pass

    T.c_float = c_float
    T.OSQPCscMatrix = OSQPCscMatrix
    T.OSQPSettings = OSQPSettings
    T.OSQPInfo = OSQPInfo
    T.OSQPSolution = OSQPSolution
    T.OSQPSolver = OSQPSolver
    return T


TYPES_F64 = make_types(C.c_double)
TYPES_F32 = make_types(C.c_float)


def bind_public_api(lib, T):
    """Declare argtypes/restypes of the public API entry points
    (include/public/osqp_api_functions.h:211-454)."""
    P = C.POINTER
    fp = P(T.c_float)
    ip = P(c_int)
    lib.osqp_set_default_settings.argtypes = [P(T.OSQPSettings)]
    lib.osqp_set_default_settings.restype = None
    lib.osqp_setup.argtypes = [P(P(T.OSQPSolver)), P(T.OSQPCscMatrix), fp,
                               P(T.OSQPCscMatrix), fp, fp, c_int, c_int,
                               P(T.OSQPSettings)]
    lib.osqp_setup.restype = c_int
    lib.osqp_solve.argtypes = [P(T.OSQPSolver)]
    lib.osqp_solve.restype = c_int
    lib.osqp_cleanup.argtypes = [P(T.OSQPSolver)]
    lib.osqp_cleanup.restype = c_int
    lib.osqp_get_solution.argtypes = [P(T.OSQPSolver), P(T.OSQPSolution)]
    lib.osqp_get_solution.restype = c_int
    lib.osqp_warm_start.argtypes = [P(T.OSQPSolver), fp, fp]
    lib.osqp_warm_start.restype = c_int
    lib.osqp_cold_start.argtypes = [P(T.OSQPSolver)]
    lib.osqp_cold_start.restype = None
    lib.osqp_update_data_vec.argtypes = [P(T.OSQPSolver), fp, fp, fp]
    lib.osqp_update_data_vec.restype = c_int
    lib.osqp_update_data_mat.argtypes = [P(T.OSQPSolver), fp, ip, c_int,
                                         fp, ip, c_int]
    lib.osqp_update_data_mat.restype = c_int
    lib.osqp_update_settings.argtypes = [P(T.OSQPSolver), P(T.OSQPSettings)]
    lib.osqp_update_settings.restype = c_int
    lib.osqp_update_rho.argtypes = [P(T.OSQPSolver), T.c_float]
    lib.osqp_update_rho.restype = c_int
    lib.osqp_capabilities.argtypes = []
    lib.osqp_capabilities.restype = c_int
    lib.osqp_version.argtypes = []
    lib.osqp_version.restype = C.c_char_p
    return lib
