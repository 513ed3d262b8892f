"""HParams work-alike for `tf.contrib.training.HParams` as used by the reference
(video_prediction/models/base_model.py:99-109): attribute access, `values()`,
`override_from_dict`, `parse("a=1,b=[2,3]")`, `set_hparam`.  Unknown keys raise ValueError, values
are cast to the type of the default (tuples/lists element-wise)."""
from __future__ import annotations

import re


def _cast_like(default, value, name):
    if isinstance(default, bool):
        if isinstance(value, str):
            if value.lower() in ('true', '1'):
                return True
            if value.lower() in ('false', '0'):
                return False
            raise ValueError('Could not parse hparam %s=%r as bool' % (name, value))
        return bool(value)
    if isinstance(default, int) and not isinstance(default, bool):
        if isinstance(value, float) and value != int(value):
            raise ValueError('hparam %s expects an int, got %r' % (name, value))
        return int(value)
    if isinstance(default, float):
        return float(value)
    if isinstance(default, str):
        return str(value)
    if isinstance(default, (list, tuple)):
        if not isinstance(value, (list, tuple)):
            value = [value]
        proto = default[0] if len(default) else None
        out = [(_cast_like(proto, v, name) if proto is not None else v) for v in value]
        return type(default)(out) if isinstance(default, tuple) else out
    return value


class HParams(object):
    def __init__(self, **kwargs):
        object.__setattr__(self, '_values', {})
        for k, v in kwargs.items():
            self.add_hparam(k, v)

    def add_hparam(self, name, value):
        if name in self._values:
            raise ValueError('Hyperparameter name is reserved: %s' % name)
        self._values[name] = value

    def set_hparam(self, name, value):
        if name not in self._values:
            raise ValueError('Unknown hyperparameter: %s' % name)
        self._values[name] = _cast_like(self._values[name], value, name)

    def override_from_dict(self, values_dict):
        for k, v in values_dict.items():
            self.set_hparam(k, v)
        return self

    def parse(self, values):
        """Parses 'name=value,name2=[v1,v2],...' and overrides."""
        pos = 0
        pat = re.compile(r'\s*(?P<name>[a-zA-Z_]\w*)\s*=\s*(?:\[(?P<list>[^\]]*)\]|(?P<val>[^,\[]*))\s*(?:,|$)')
        values = values.strip()
        while pos < len(values):
            m = pat.match(values, pos)
            if not m:
                raise ValueError('Malformed hyperparameter value: %s' % values[pos:])
            pos = m.end()
            name = m.group('name')
            if m.group('list') is not None:
                items = [s.strip() for s in m.group('list').split(',') if s.strip() != '']
                val = [self._parse_scalar(s) for s in items]
            else:
                val = self._parse_scalar(m.group('val').strip())
            self.set_hparam(name, val)
        return self

    @staticmethod
    def _parse_scalar(s):
        try:
            return int(s)
        except ValueError:
            pass
        try:
            return float(s)
        except ValueError:
            pass
        return s

    def values(self):
        return dict(self._values)

    def get(self, key, default=None):
        return self._values.get(key, default)

    def __contains__(self, key):
        return key in self._values

    def __getattr__(self, name):
        try:
            return object.__getattribute__(self, '_values')[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in self._values:
            self.set_hparam(name, value)
        else:
            object.__setattr__(self, name, value)

    def __repr__(self):
        return 'HParams(%s)' % ', '.join('%s=%r' % kv for kv in sorted(self._values.items()))
