"""String enums of the public API.

Host-side constants only; they mirror the values of the reference's
``torchsde/settings.py:16-61`` (METHODS :29-38, NOISE_TYPES :41-45, SDE_TYPES :48-50,
LEVY_AREA_APPROXIMATIONS :53-57, METHOD_OPTIONS :60-61) so that user code written against the
reference keeps working unchanged.
"""


class _Enum(type):
    def all(cls):
        return sorted(v for k, v in vars(cls).items() if not k.startswith('_') and isinstance(v, str))

    def __contains__(cls, item):
        return item in cls.all()

    def __str__(cls):
        return str(cls.all())

    def __iter__(cls):
        return iter(cls.all())


class METHODS(metaclass=_Enum):
    euler = 'euler'
    milstein = 'milstein'
    srk = 'srk'
    midpoint = 'midpoint'
    reversible_heun = 'reversible_heun'
    adjoint_reversible_heun = 'adjoint_reversible_heun'
    heun = 'heun'
    log_ode_midpoint = 'log_ode'
    euler_heun = 'euler_heun'


class NOISE_TYPES(metaclass=_Enum):  # noqa
    general = 'general'
    diagonal = 'diagonal'
    scalar = 'scalar'
    additive = 'additive'


class SDE_TYPES(metaclass=_Enum):  # noqa
    ito = 'ito'
    stratonovich = 'stratonovich'


class LEVY_AREA_APPROXIMATIONS(metaclass=_Enum):  # noqa
    none = 'none'
    space_time = 'space-time'
    davie = 'davie'
    foster = 'foster'


class METHOD_OPTIONS(metaclass=_Enum):  # noqa
    grad_free = 'grad_free'
