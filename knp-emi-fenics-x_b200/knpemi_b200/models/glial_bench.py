"""Passive glial membrane of the single-cell benchmark case (mV, ms).

Builtin restatement of examples/benchmark/mm_glial.py (state :11, parameters
:46-62, right-hand side :120-204).  21 parameters in an order that differs
from ``glial_tissue``; ``K_e_init`` / ``K_i_init`` are literals, the Kir
reference potential uses ``1/psi`` (:176) and the Kir constants are
18.4/42.4 and 18.5/42.5 (:178-181).
"""
import math

import numpy as np

from ._protocol import rhs_cfunc, table_functions

STATES = (("V", -85.85765274084892),)

PARAMETERS = (
    ("psi", 0.0),
    ("g_leak_Cl", 0.05), ("g_leak_Na", 0.1), ("g_leak_K", 1.696),
    ("z_Na", 0.0), ("z_K", 0.0), ("z_Cl", 0.0),
    ("Cm", 0.0), ("stim_amplitude", 0.0),
    ("I_ch_Na", 0.0), ("I_ch_K", 0.0), ("I_ch_Cl", 0.0),
    ("K_e", 0.0), ("K_i", 0.0), ("Na_e", 0.0), ("Na_i", 0.0),
    ("Cl_e", 0.0), ("Cl_i", 0.0),
    ("m_K", 1.5), ("m_Na", 10.0), ("I_max", 10.75975),
)

(init_state_values, init_parameter_values,
 state_indices, parameter_indices) = table_functions(STATES, PARAMETERS)


@rhs_cfunc
def rhs_numba(t, states, values, parameters):
    psi = parameters[0]
    g_leak_Cl = parameters[1]
    g_leak_Na = parameters[2]
    g_leak_K = parameters[3]
    z_K = parameters[5]
    z_Cl = parameters[6]
    Cm = parameters[7]
    K_e = parameters[12]
    K_i = parameters[13]
    Na_e = parameters[14]
    Na_i = parameters[15]
    Cl_e = parameters[16]
    Cl_i = parameters[17]
    m_K = parameters[18]
    m_Na = parameters[19]
    I_max = parameters[20]

    V = states[0]

    E_Na = 1/psi/z_K * math.log(Na_e/Na_i)
    E_K = 1/psi/z_K * math.log(K_e/K_i)
    E_Cl = 1/psi/z_Cl * math.log(Cl_e/Cl_i)

    K_e_init = 3.092970607490389
    K_i_init = 99.3100014897692

    i_pump = I_max*(K_e/(K_e + m_K))*(Na_i**(1.5)/(Na_i**(1.5) + m_Na**(1.5)))

    # inward-rectifying K conductance
    E_K_init = 1/psi*np.log(K_e_init/K_i_init)
    dphi = V - E_K
    A = 1 + np.exp(18.4/42.4)
    B = 1 + np.exp(-(0.1186e3 + E_K_init)/0.0441e3)
    C = 1 + np.exp((dphi + 0.0185e3)/0.0425e3)
    D = 1 + np.exp(-(0.1186e3 + V)/0.0441e3)
    g_Kir = np.sqrt(K_e/K_e_init)*(A*B)/(C*D)

    i_Kir = g_leak_K*g_Kir*(V - E_K)
    i_Na = g_leak_Na*(V - E_Na) + 3*i_pump
    i_K = i_Kir - 2*i_pump
    i_Cl = g_leak_Cl*(V - E_Cl)

    parameters[9] = i_Na
    parameters[10] = i_K
    parameters[11] = i_Cl

    values[0] = (-i_K - i_Na - i_Cl)/Cm
