"""Manufactured solution of Solver/test/NavierStokes/Convergence (SETUP/ProblemFile.f90:31-133, state_source_in_point):
exact state Q, source S and time derivative QDot at points (x, y, z) and time t.  Test input, not product code."""
import numpy as np

PI = np.pi
# Gauss-Lobatto weights the reference's FinalCheck uses for its error norms (ProblemFile.f90:657-664), as printed there
W_LGL7 = np.array([0.035714285714286, 0.210704227143506, 0.341122692483504, 0.412458794658704,
                   0.412458794658704, 0.341122692483504, 0.210704227143506, 0.035714285714286])


def state_source_in_point(x, y, z, t, gm1, gMa2, mu, kap):
    c_u, c_v = 1.0, 2.0
    c_w = -c_u - c_v
    S_suth = 0.381923076923077
    arg = PI * (x + y + z - 2.0 * t)
    dargdx, dargdt = PI, -2.0 * PI
    rho = 2.0 + 0.1 * np.sin(arg)
    u = c_u * rho - 1.0
    v = c_v * rho - 4.0
    w = c_w * rho + 5.0
    e = rho
    vtot = u * u + v * v + w * w
    p = gm1 * (rho * e - 0.5 * rho * vtot)
    Temp = gMa2 * gm1 * (e - 0.5 * vtot)
    suther = (1.0 + S_suth) / (Temp + S_suth) * Temp * np.sqrt(Temp)
    dsutherdT = 1.5 * suther / Temp - suther / (S_suth + Temp)
    Q = np.stack([rho, rho * u, rho * v, rho * w, rho * e], axis=-1)
    rho_t = 0.1 * dargdt * np.cos(arg)
    u_t, v_t, w_t, e_t = c_u * rho_t, c_v * rho_t, c_w * rho_t, rho_t
    QDot = np.stack([rho_t, rho_t * u + rho * u_t, rho_t * v + rho * v_t, rho_t * w + rho * w_t, rho_t * e + rho * e_t], axis=-1)
    # gradients: every Cartesian component is the same (arg depends on x + y + z)
    rho_x = 0.1 * dargdx * np.cos(arg)
    u_x, v_x, w_x, e_x = c_u * rho_x, c_v * rho_x, c_w * rho_x, rho_x
    p_x = gm1 * rho_x * (e - 0.5 * vtot) + gm1 * rho * (e_x - u * u_x - v * v_x - w * w_x)
    Temp_x = gMa2 * gm1 * (e_x - u * u_x - v * v_x - w * w_x)
    div_v = u_x + v_x + w_x
    gu = [u_x, v_x, w_x]                       # grad_u(i, j) = d u_j / d x_i = gu[j] for every i
    tau = [[mu * suther * (gu[j] + gu[i] - (2.0 * div_v * (1.0 if i == j else 0.0) / 3.0)) for j in range(3)] for i in range(3)]
    rho_xx = -0.1 * dargdx ** 2 * np.sin(arg)
    u_xx, v_xx, w_xx, e_xx = c_u * rho_xx, c_v * rho_xx, c_w * rho_xx, rho_xx
    Temp_xx = gMa2 * gm1 * (e_xx - u_x * u_x - v_x * v_x - w_x * w_x - u * u_xx - v * v_xx - w * w_xx)   # every entry
    lap = [u_xx + u_xx + u_xx, v_xx + v_xx + v_xx, w_xx + w_xx + w_xx]
    mixed = u_xx + v_xx + w_xx
    div_tau = [mu * suther * (mixed + lap[d] - (2.0 / 3.0) * mixed) for d in range(3)]
    div_tau = [div_tau[i] + (dsutherdT / suther) * (tau[i][0] * Temp_x + tau[i][1] * Temp_x + tau[i][2] * Temp_x) for i in range(3)]
    sum_gt = 0.0
    for j in range(3):                        # sum(grad_u*tau), column-major
        for i in range(3):
            sum_gt = sum_gt + gu[j] * tau[i][j]
    visc_work = sum_gt + u * div_tau[0] + v * div_tau[1] + w * div_tau[2]
    heat_flux = kap * dsutherdT * (Temp_x ** 2 + Temp_x ** 2 + Temp_x ** 2) + kap * suther * (Temp_xx + Temp_xx + Temp_xx)
    S0 = rho_t + u * rho_x + v * rho_x + w * rho_x + rho * div_v
    S1 = u * S0 + rho * u_t + rho * u * u_x + rho * v * u_x + rho * w * u_x + p_x - div_tau[0]
    S2 = v * S0 + rho * v_t + rho * u * v_x + rho * v * v_x + rho * w * v_x + p_x - div_tau[1]
    S3 = w * S0 + rho * w_t + rho * u * w_x + rho * v * w_x + rho * w * w_x + p_x - div_tau[2]
    S4 = e * S0 + rho * e_t + rho * (u * e_x + v * e_x + w * e_x) + p * div_v + u * p_x + v * p_x + w * p_x - visc_work - heat_flux
    S = np.stack([S0, S1, S2, S3, S4], axis=-1)
    return Q, S, QDot
