"""GPU twins of troute.routing.fast_reach.reach (reach.pyx): the python-callable single-segment and
single-reach kernels the reference's kernel tests use (mc_sseg_stime_NOLOOP_demo.py:434-454).

Both run on the device through the C ABI (trt_mc_segment_batch); there is no CPU fallback.
"""
import numpy as np

from ...network import mc_segment_batch


def compute_reach_kernel(dt, qup, quc, qdp, ql, dx, bw, tw, twcc, n, ncc, cs, s0, velp, depthp, device=0):
    """reach.pyx:66-103 -> dict(qdc, velc, depthc, cn, ck, X) (reach.pxd:1-7), float32."""
    row = np.array([[dt, qup, quc, qdp, ql, dx, bw, tw, twcc, n, ncc, cs, s0, velp, depthp]], dtype=np.float32)
    o = mc_segment_batch(row, device=device)[0]
    return {"qdc": o[0], "velc": o[1], "depthc": o[2], "ck": o[3], "cn": o[4], "X": o[5]}


def boundary_shape():
    return 2


def previous_state_cols():
    return 3


def parameter_inputs_cols():
    return 13


def output_buffer_cols():
    return 3


def compute_reach(boundary, previous_state, parameter_inputs, output_buffer, size=0, device=0):
    """reach.pyx:119-206: walk one reach for one timestep.

    boundary          [qup, quc] entering the head segment                                  (:166-167)
    previous_state    [n, 3]  (qdp, velp, depthp) per segment                              (:181-183)
    parameter_inputs  [n, >=10] (qlat, dt, dx, bw, tw, twcc, n, ncc, cs, s0) per segment    (:170-179)
    output_buffer     [n, 3]  receives (qdc, velc, depthc)                                  (:202-204)
    A segment's upstream flows are its predecessor's: quc = qdc(i-1), qup = qdp(i-1) (:202, :206); the chain is a
    data dependence, so the segments are solved one launch after another.  Same checks and exceptions as :141-163."""
    boundary = np.asarray(boundary, dtype=np.float32)
    previous_state = np.asarray(previous_state, dtype=np.float32)
    parameter_inputs = np.asarray(parameter_inputs, dtype=np.float32)
    if size > 0:
        rows = int(size)
        if parameter_inputs.shape[0] < rows or output_buffer.shape[0] < rows or previous_state.shape[0] < rows:
            raise ValueError(f"axis 0 is not long enough for {size}")
    else:
        rows = previous_state.shape[0]
        if rows != parameter_inputs.shape[0] or rows != output_buffer.shape[0]:
            raise ValueError("axis 0 of input arguments do not agree")
    if boundary.shape[0] < 2 or parameter_inputs.shape[1] < 10 or output_buffer.shape[1] < 3 \
            or previous_state.shape[1] < 3:
        raise IndexError
    qup, quc = boundary[0], boundary[1]
    for i in range(rows):
        p = parameter_inputs[i]
        qdp, velp, depthp = previous_state[i, 0], previous_state[i, 1], previous_state[i, 2]
        row = np.array([[p[1], qup, quc, qdp, p[0], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9], velp, depthp]],
                       dtype=np.float32)
        o = mc_segment_batch(row, device=device)[0]
        output_buffer[i, 0] = quc = o[0]
        output_buffer[i, 1] = o[1]
        output_buffer[i, 2] = o[2]
        qup = qdp
    return output_buffer
