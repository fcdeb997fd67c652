"""ctypes mirrors of the descriptor structs in csrc/ops.h (field order and types must match exactly;
`_lib.py` checks every sizeof against the library's ds_sizeof())."""
import ctypes as C

P = C.c_uint64      # pointer fields: absolute device address or a reference (space << 60 | offset)
I32 = C.c_int32
I64 = C.c_int64
F32 = C.c_float

DS_OP_GEMM, DS_OP_GN_STATS, DS_OP_GN_APPLY, DS_OP_SOFTMAX, DS_OP_POSEMB, DS_OP_LINEAR = 1, 2, 3, 4, 5, 6
DS_OP_PREP_INPUT, DS_OP_CHANMEAN, DS_OP_MEMSET, DS_OP_LAYERNORM, DS_OP_GEGLU, DS_OP_GN_FINALIZE, DS_OP_ATTN, DS_OP_EMBED = 7, 8, 9, 10, 11, 12, 13, 14
DS_IO_X, DS_IO_D, DS_IO_SIGMA, DS_IO_LABELS, DS_IO_BOTTLENECK, DS_IO_CTX = 0, 1, 2, 3, 4, 5
DS_M_X0, DS_M_EPS, DS_M_DIV, DS_M_NONE = 0, 1, 2, 3
DS_F8_SH_A16, DS_F8_SH_LO8, DS_F8_SH_HI8 = 6, 13, 2      # csrc/ops.h: power-of-two operand scales of the f8 GEMM mode

SPACE_ABS, SPACE_ARENA, SPACE_WEIGHTS, SPACE_IO = 0, 1, 2, 3


def ref(space, offset):
    return (space << 60) | int(offset)


class GemmDesc(C.Structure):
    _fields_ = [
        ('a_ptr', P), ('a_dims', I64 * 4), ('a_strides', I64 * 3), ('a_box', I32 * 4), ('a_plane_n', I32),
        ('a2_ptr', P), ('a2_c', I64), ('a2_plane_n', I32), ('nkb_aux', I32),
        ('b_ptr', P), ('b_dims', I64 * 3), ('b_strides', I64 * 2), ('b_plane_batch', I32),
        ('BN', I32), ('m_tiles', I32), ('n_tiles', I32), ('num_z', I32), ('nh', I32),
        ('taps', I32), ('cpb', I32), ('npass', I32), ('a_mode', I32), ('conv_H', I32), ('conv_W', I32),
        ('a_c_per_zh', I32), ('a_n_per_zb', I32), ('a_n_per_zh', I32),
        ('b_k0', I32), ('b_k_per_zh', I32), ('b_row_per_zh', I32), ('b_z_per_zb', I32), ('b_z_per_zh', I32),
        ('m_valid', I32), ('n_valid', I32),
        ('out_f32', P), ('out_h16', P), ('o_zb', I64), ('o_zh', I64), ('ldo', I64), ('o_plane', I64),
        ('bias_n', P), ('bias_m', P), ('rowvec', P), ('rowvec_stride', I64), ('rows_per_sample', I32), ('f8', I32),
        ('residual', P), ('ldr', I64), ('scale', F32),
        ('edm_out', I32), ('edm_x', P), ('edm_coef', P), ('edm_coef_stride', I32), ('edm_C', I32), ('edm_D', P),
        ('st_quads', P),
        ('tap_dh', I32 * 9), ('tap_dw', I32 * 9), ('tap_cb', I32 * 9), ('acc_scale', F32), ('st_unit', I32),
    ]


class GnStatsDesc(C.Structure):
    _fields_ = [('src0', P), ('src1', P), ('C0', I32), ('C1', I32), ('HW', I32), ('B', I32), ('groups', I32), ('pad0', I32),
                ('sums', P)]


class GnApplyDesc(C.Structure):
    _fields_ = [('src0', P), ('src1', P), ('C0', I32), ('C1', I32), ('H', I32), ('W', I32), ('B', I32), ('groups', I32),
                ('sums', P), ('gamma', P), ('beta', P), ('eps', F32), ('silu', I32), ('ada', P), ('ada_stride', I64),
                ('resample', I32), ('nplanes', I32), ('out_act', P), ('out_raw', P), ('out_raw_f32', P), ('fmt', I32), ('pad0', I32), ('coef', P)]


class SoftmaxDesc(C.Structure):
    _fields_ = [('S', P), ('P', P), ('rows', I64), ('L', I32), ('nplanes', I32), ('pitch_in', I32), ('pitch_out', I32)]


class PosembDesc(C.Structure):
    _fields_ = [('sigma', P), ('nsig', I32), ('num_channels', I32), ('endpoint', I32), ('swap_sincos', I32),
                ('sigma_data', F32), ('mode', I32), ('coef', P), ('emb', P)]


class LinearDesc(C.Structure):
    _fields_ = [('in_', P), ('in_stride', I64), ('W', P), ('b', P), ('add', P), ('add_stride', I64), ('out', P),
                ('n_rows', I32), ('in_f', I32), ('out_f', I32), ('act', I32), ('in_scale', F32), ('pad0', I32)]


class PrepInputDesc(C.Structure):
    _fields_ = [('x', P), ('coef', P), ('coef_stride', I32), ('B', I32), ('C', I32), ('HW', I32), ('nplanes', I32),
                ('x_batch', I32), ('out', P)]


class ChanmeanDesc(C.Structure):
    _fields_ = [('src', P), ('out', P), ('rows', I64), ('C', I32), ('pad0', I32)]


class LayernormDesc(C.Structure):
    _fields_ = [('src', P), ('gamma', P), ('beta', P), ('out', P), ('rows', I64), ('C', I32), ('nplanes', I32), ('eps', F32), ('fmt', I32)]


class GegluDesc(C.Structure):
    _fields_ = [('src', P), ('out', P), ('rows', I64), ('I', I32), ('nplanes', I32), ('fmt', I32), ('mode', I32)]


class GnFinalizeDesc(C.Structure):
    _fields_ = [('quads0', P), ('quads1', P), ('C0', I32), ('C1', I32), ('slabs_per_sample', I32), ('B', I32), ('groups', I32),
                ('unit0', I32), ('sums', P), ('gamma', P), ('beta', P), ('ada', P), ('ada_stride', I64), ('eps', F32), ('HW', I32), ('coef', P),
                ('unit1', I32), ('pad1', I32)]


class AttnDesc(C.Structure):
    _fields_ = [('q', P), ('k', P), ('vt', P), ('out', P), ('B', I32), ('nh', I32), ('L', I32), ('Lk', I32),
                ('q_pitch', I32), ('q_c0', I32), ('k_pitch', I32), ('k_c0', I32), ('vt_pitch', I32), ('o_pitch', I32),
                ('nplanes', I32), ('scale', F32), ('causal', I32), ('pad0', I32)]


class EmbedDesc(C.Structure):
    _fields_ = [('ids', P), ('tok', P), ('pos', P), ('out', P), ('rows', I64), ('T', I32), ('C', I32), ('vocab', I32), ('pad0', I32)]


class MemsetDesc(C.Structure):
    _fields_ = [('ptr', P), ('bytes', I64)]


class _OpUnion(C.Union):
    _fields_ = [('gemm', GemmDesc), ('gn_stats', GnStatsDesc), ('gn_apply', GnApplyDesc), ('softmax', SoftmaxDesc),
                ('posemb', PosembDesc), ('linear', LinearDesc), ('prep_input', PrepInputDesc), ('chanmean', ChanmeanDesc),
                ('memset', MemsetDesc), ('layernorm', LayernormDesc), ('geglu', GegluDesc), ('gn_finalize', GnFinalizeDesc), ('attn', AttnDesc),
                ('embed', EmbedDesc)]


class PlanOp(C.Structure):
    _fields_ = [('type', I32), ('tag', I32), ('u', _OpUnion)]


SIZEOF_CHECKS = {
    0: PlanOp, DS_OP_GEMM: GemmDesc, DS_OP_GN_STATS: GnStatsDesc, DS_OP_GN_APPLY: GnApplyDesc, DS_OP_SOFTMAX: SoftmaxDesc,
    DS_OP_POSEMB: PosembDesc, DS_OP_LINEAR: LinearDesc, DS_OP_PREP_INPUT: PrepInputDesc, DS_OP_CHANMEAN: ChanmeanDesc,
    DS_OP_MEMSET: MemsetDesc, DS_OP_LAYERNORM: LayernormDesc, DS_OP_GEGLU: GegluDesc, DS_OP_GN_FINALIZE: GnFinalizeDesc, DS_OP_ATTN: AttnDesc,
    DS_OP_EMBED: EmbedDesc,
}

UNION_FIELD = {
    DS_OP_GEMM: 'gemm', DS_OP_GN_STATS: 'gn_stats', DS_OP_GN_APPLY: 'gn_apply', DS_OP_SOFTMAX: 'softmax', DS_OP_POSEMB: 'posemb',
    DS_OP_LINEAR: 'linear', DS_OP_PREP_INPUT: 'prep_input', DS_OP_CHANMEAN: 'chanmean', DS_OP_MEMSET: 'memset',
    DS_OP_LAYERNORM: 'layernorm', DS_OP_GEGLU: 'geglu', DS_OP_GN_FINALIZE: 'gn_finalize', DS_OP_ATTN: 'attn', DS_OP_EMBED: 'embed',
}
OP_TYPE_OF = {GemmDesc: DS_OP_GEMM, GnStatsDesc: DS_OP_GN_STATS, GnApplyDesc: DS_OP_GN_APPLY, SoftmaxDesc: DS_OP_SOFTMAX,
              PosembDesc: DS_OP_POSEMB, LinearDesc: DS_OP_LINEAR, PrepInputDesc: DS_OP_PREP_INPUT, ChanmeanDesc: DS_OP_CHANMEAN,
              MemsetDesc: DS_OP_MEMSET, LayernormDesc: DS_OP_LAYERNORM, GegluDesc: DS_OP_GEGLU, GnFinalizeDesc: DS_OP_GN_FINALIZE, AttnDesc: DS_OP_ATTN,
              EmbedDesc: DS_OP_EMBED}
