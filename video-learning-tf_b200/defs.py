"""String constants of the configuration language (`defs.x.y` values in the YAML files).

Mirrors the vocabulary of the reference's defs_.py:4-122 so that existing configuration files keep their
meaning; `check` validates a dotted name against an expected family exactly like defs_.py:6-34 (same error texts).
"""
from .utils import error


class _Family(object):
    """A namespace of string constants: attribute name == value."""

    def __init__(self, name, *values):
        self._name = name
        self._values = list(values)
        for v in values:
            setattr(self, v, v)

    def __contains__(self, v):
        return v in self._values

    def __iter__(self):
        return iter(self._values)

    def __repr__(self):
        return "<defs.%s>" % self._name


class _Defs(object):
    def __init__(self):
        fam = {
            "representation": ("dcnn", "fc", "nop"),
            "classifier": ("fc", "lstm"),
            "phase": ("train", "val"),
            "input_mode": ("video", "image", "vectors"),
            "net_input": ("visual", "labels"),
            "dataset_tag": ("main", "aux"),
            "data_format": ("raw", "tfrecord", "synthetic", "npy"),  # synthetic / npy: TF-free feeds of this framework
            "fusion_method": ("avg", "last", "concat", "reshape", "state", "ibias", "maximum"),
            "fusion_type": ("early", "late", "none", "main", "aux"),
            "clipframe_mode": ("rand_frames", "rand_clips", "iterative"),
            "generation_error": ("abort", "compromise", "report"),
            "batch_item": ("default", "clip"),
            "optim": ("sgd", "rmsprop", "adam"),
            "decay": ("exp", "staircase"),
            "periodicity": ("interval", "drops"),
            "label_type": ("single", "multiple"),
            "imgproc": ("rand_mirror", "rand_crop", "center_crop", "resize", "raw_resize", "sub_mean"),
        }
        self._families = {}
        for name, values in fam.items():
            f = _Family(name, *values)
            self._families[name] = f
            setattr(self, name, f)

        class names:
            global_step, latest_savefile = "global_step", "latest"
        self.names = names

    def check(self, arg, should_belong_to, do_boolean=False):
        """Resolve "defs.family.value" and require it to live under `should_belong_to` (a family or a tuple of
        families).  Returns the value; with do_boolean returns (ok, value) instead of raising."""
        allowed = should_belong_to if isinstance(should_belong_to, (tuple, list)) else (should_belong_to,)

        def fail(msg):
            if do_boolean:
                return (False, None)
            error(msg)

        if not isinstance(arg, str):
            return fail("Invalid def : %s" % str(arg))
        parts = arg.split(".")
        if parts[0] != "defs":
            return fail("Invalid def : %s" % arg)
        if len(parts) != 3:
            return fail("Parameter [%s] is not defined for [%s]" % (arg, self))
        fam = self._families.get(parts[1])
        if fam is None:
            return fail("Parameter [%s] is not defined for [%s]" % (parts[1], self))
        if parts[2] not in fam:
            return fail("Parameter [%s] is not defined for [%s]" % (parts[2], fam))
        if not any(fam is a for a in allowed):
            return fail("Supplied parameter [%s] should be a child of def [%s]" % (arg, should_belong_to))
        return (True, parts[2]) if do_boolean else parts[2]


defs = _Defs()
