"""The literal problem instances and golden vectors of the reference's test file
(/root/reference/test/runtests.jl), restated as data so that the tests can run
where /root/reference does not exist (the GPU box).  Line numbers cite that file."""
import numpy as np

POC, SOC = 0, 1

# --- "Vector operations", test/runtests.jl:10-28
VEC_TV1 = np.array([1, 1, 1, 1, 2, 3], dtype=np.float64)
VEC_TV2 = np.array([1, 1, 1, 1, 5, 6], dtype=np.float64)
VEC_CONES = ((POC, 0, 3), (SOC, 3, 3))

# --- "Nesterov-Todd Scalings", test/runtests.jl:30-48
NT_S = np.array([1, 1, 1, 9, 2, 3], dtype=np.float64)
NT_Z = np.array([1, 1, 1, 22, 5, 6], dtype=np.float64)
NT_CONES = ((POC, 0, 3), (SOC, 3, 3))

# --- "Squared NT Scalings", test/runtests.jl:63-90
SQ_G = np.array([[0, 0, 1.0], [0, 0, -1], [0, -1, 0], [-1, 0, 0]])
SQ_CONES = ((POC, 0, 1), (SOC, 1, 3))
SQ_U1 = np.array([3.414213562373095, 2.414213562373095, 1.0, 1.0])
SQ_V1 = np.array([1.414213562373095, 2.414213562373095, -1.0, -1.0])
SQ_U2 = np.array([2.5571536140045033, 2.594521365784194, 1.6419234525608073, 1.6419234525608069])
SQ_V2 = np.array([0.42421356237309515, 1.424213562373095, -1.0, -0.9999999999999997])

# --- "KKT reference solution", test/runtests.jl:95-128
KKT = dict(
    c=np.array([-1.0, -1.0, 1.0]),
    A=np.zeros((0, 3)), b=np.zeros(0),
    G=np.array([[0, 0, 1.0], [0, 0, -1], [0, -1, 0], [-1, 0, 0]]),
    h=np.array([5.0, 0.0, 0.0, 0.0]),
    cones=((POC, 0, 1), (SOC, 1, 3)),
    x=np.array([3.5093936289670493, 3.5093936289670546, 4.984484369850202]),
    y=np.zeros(0),
    z=np.array([0.4416936891219317, 1.4416936891219376, -1.0000000000000004, -0.9999999999999993]),
    s=np.array([0.025571536140045037, 4.994540275840449, 3.5093936289670546, 3.5093936289670484]),
    dx=np.array([6.661338147750939e-16, -4.440892098500626e-16, 5.995204332975845e-15]),
    dy=np.zeros(0),
    dz=np.array([-0.010055905990247638, -0.01005590599024675, -0.0, 8.881784197001252e-16]),
    ds=np.array([-0.011294786134211294, -0.18180993781041443, -0.06493037168608469, -0.06493037168608427]),
    cxr=np.array([0.02479569536244916, 0.02479569536250157, 0.013918468357446631]),
    cyr=np.zeros(0),
    czr=np.array([-0.02758755986830461, -0.02758755986827908, 1.3092496892856528e-14, 3.9336377471191126e-14]),
    csr=np.array([-0.02397437434769427, 0.003862562367179831, 0.024795695362486953, 0.02479569536243503]),
)


def socp1():
    """test/runtests.jl:130-146; x* ~ [3.53553, 3.53553, 5.0]."""
    return dict(c=np.array([-1.0, -1.0, 1.0]), A=np.zeros((0, 3)), b=np.zeros(0),
                G=np.array([[0, 0, 1.0], [0, 0, -1], [0, -1, 0], [-1, 0, 0]]),
                h=np.array([5.0, 0.0, 0.0, 0.0]), cones=((POC, 0, 1), (SOC, 1, 3)),
                xstar=np.array([3.53553, 3.53553, 5.0]))


_G2 = np.array([[12.0, 6.0, -5.0], [13.0, -3.0, -5.0], [12.0, -12.0, 6.0], [3.0, -6.0, 10.0],
                [3.0, -6.0, -2.0], [-1.0, -9.0, -2.0], [1.0, 19.0, -3.0]])
_H2 = np.array([-12.0, -3.0, -2.0, 27.0, 0.0, 3.0, -42.0])


def socp2():
    """test/runtests.jl:148-167; x* ~ [-5.01467, -5.7669, -8.52176]."""
    return dict(c=np.array([-2.0, 1.0, 5.0]), A=np.zeros((0, 3)), b=np.zeros(0), G=_G2.copy(), h=_H2.copy(),
                cones=((SOC, 0, 3), (SOC, 3, 4)), xstar=np.array([-5.01467, -5.7669, -8.52176]))


def socp3():
    """test/runtests.jl:169-191; one equality x1 = -3; x* ~ [-3.0, -4.82569, -6.64011]."""
    return dict(c=np.array([-2.0, 1.0, 5.0]), A=np.array([[1.0, 0.0, 0.0]]), b=np.array([-3.0]),
                G=_G2.copy(), h=_H2.copy(), cones=((SOC, 0, 3), (SOC, 3, 4)),
                xstar=np.array([-3.0, -4.82569, -6.64011]))


def control(n=50):
    """"Linear optimal control", test/runtests.jl:204-244 (1-based indices shifted
    to 0-based).  3n variables, 2n+2 equalities, one SOC(0,n); sing = true."""
    c = np.zeros(3 * n)
    c[3 * n - 1] = 1.0
    vel = list(range(0, n))
    pos = list(range(n, 2 * n))
    force = list(range(2 * n, 3 * n - 1))
    A = np.zeros((2 * n + 2, 3 * n))
    b = np.zeros(2 * n + 2)
    A[vel[0], vel[0]] = 1.0
    b[vel[0]] = 1.0
    A[pos[0], pos[0]] = 1.0
    b[pos[0]] = 0.0
    for stp in range(1, n):
        A[vel[stp], vel[stp]] = -1.0
        A[vel[stp], vel[stp - 1]] = 1.0
        A[vel[stp], force[stp - 1]] = 1.0
        A[pos[stp], pos[stp]] = -1.0
        A[pos[stp], pos[stp - 1]] = 1.0
        A[pos[stp], vel[stp - 1]] = 1.0
    A[2 * n, vel[n - 1]] = 1.0
    A[2 * n + 1, pos[n - 1]] = 1.0
    G = np.zeros((n, 3 * n))
    G[0, 3 * n - 1] = -1.0
    for i in range(1, n):
        G[i, 2 * n + i - 1] = -1.0
    h = np.zeros(n)
    return dict(c=c, A=A, b=b, G=G, h=h, cones=((SOC, 0, n),), xstar=None)


ALL_C1 = dict(socp1=socp1, socp2=socp2, socp3=socp3, control=control)
