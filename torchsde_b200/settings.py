"""String constants of the public API.

Host-side names only.  The values are the strings users of the reference already pass
(``method='srk'``, ``noise_type='diagonal'``, ``levy_area_approximation='space-time'`` ...; reference
``torchsde/settings.py:29-61``), exposed as attribute containers that also support ``in``, iteration,
``str()`` and ``.all()`` (sorted values) like the reference's ``ContainerMeta`` (:16-25).
"""


class _Names(type):
    def all(cls):
        return sorted(cls._values)

    def __contains__(cls, item):
        return item in cls._values

    def __iter__(cls):
        return iter(cls.all())

    def __str__(cls):
        return str(cls.all())


def _names(name, **members):
    ns = dict(members)
    ns['_values'] = frozenset(members.values())
    return _Names(name, (), ns)


METHODS = _names('METHODS', euler='euler', milstein='milstein', srk='srk', midpoint='midpoint', heun='heun',
                 euler_heun='euler_heun', log_ode_midpoint='log_ode', reversible_heun='reversible_heun',
                 adjoint_reversible_heun='adjoint_reversible_heun')
NOISE_TYPES = _names('NOISE_TYPES', general='general', diagonal='diagonal', scalar='scalar', additive='additive')
SDE_TYPES = _names('SDE_TYPES', ito='ito', stratonovich='stratonovich')
# none: no Levy area; space-time: exact space-time Levy area; davie / foster: approximations of the Levy area
LEVY_AREA_APPROXIMATIONS = _names('LEVY_AREA_APPROXIMATIONS', none='none', space_time='space-time', davie='davie',
                                  foster='foster')
METHOD_OPTIONS = _names('METHOD_OPTIONS', grad_free='grad_free')
