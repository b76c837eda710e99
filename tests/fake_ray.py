"""A tiny in-process stand-in for the parts of the ray API the executor uses (actors run synchronously)."""
import sys
import types


class _Ref:
    def __init__(self, fn):
        self._fn, self._done, self._val = fn, False, None

    def get(self):
        if not self._done:
            self._val, self._done = self._fn(), True
        return self._val


class _Method:
    def __init__(self, obj, name):
        self.obj, self.name = obj, name

    def remote(self, *a, **k):
        ref = _Ref(lambda: getattr(self.obj, self.name)(*a, **k))
        ref.get()                      # eager, like a task that already ran
        return ref


class _Handle:
    def __init__(self, obj):
        self._obj = obj

    def __getattr__(self, name):
        return _Method(self._obj, name)


class _ActorClass:
    def __init__(self, cls):
        self.cls, self.opts = cls, {}

    def options(self, **kw):
        c = _ActorClass(self.cls)
        c.opts = kw
        return c

    def remote(self, *a, **k):
        h = _Handle(self.cls(*a, **k))
        STATE["actors"].append((h, self.opts))
        return h


STATE = {"init": False, "actors": [], "killed": 0, "pgs": []}


def install():
    ray = types.ModuleType("ray")
    ray.is_initialized = lambda: STATE["init"]
    ray.init = lambda **kw: STATE.__setitem__("init", True)
    ray.remote = lambda cls: _ActorClass(cls)
    ray.get = lambda r: [x.get() for x in r] if isinstance(r, list) else r.get()
    ray.get_gpu_ids = lambda: []
    ray.kill = lambda a: STATE.__setitem__("killed", STATE["killed"] + 1)
    util = types.ModuleType("ray.util")
    util.get_node_ip_address = lambda: "127.0.0.1"
    pgm = types.ModuleType("ray.util.placement_group")

    class PG:
        def __init__(self, bundles, strategy):
            self.bundles, self.strategy = bundles, strategy

        def ready(self):
            return _Ref(lambda: True)

    def placement_group(bundles, strategy="PACK"):
        pg = PG(bundles, strategy)
        STATE["pgs"].append(pg)
        return pg
    pgm.placement_group = placement_group
    ss = types.ModuleType("ray.util.scheduling_strategies")

    class PlacementGroupSchedulingStrategy:
        def __init__(self, placement_group, placement_group_bundle_index):
            self.pg, self.index = placement_group, placement_group_bundle_index
    ss.PlacementGroupSchedulingStrategy = PlacementGroupSchedulingStrategy
    ray.util = util
    util.placement_group = pgm
    util.scheduling_strategies = ss
    mods = {"ray": ray, "ray.util": util, "ray.util.placement_group": pgm, "ray.util.scheduling_strategies": ss}
    sys.modules.update(mods)
    return list(mods)


def uninstall(names):
    for n in names:
        sys.modules.pop(n, None)
    STATE.update(init=False, actors=[], killed=0, pgs=[])
