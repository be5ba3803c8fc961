"""Minimal circom-semantics executor used by the oracle (test infrastructure).

A template is a Python function that receives a `Comp` (the component
instance) and declares / assigns signals in the same order the circom source
does.  The executor reproduces the two compiler behaviours that shape a
witness [circom 0.5.45, un-vendored; restated, layout parity unpinned]:

* numbering: signals are numbered depth-first in the order their declarations
  execute; a sub-component's signals are laid out where its instantiation
  statement (`component x = T()`) executes -- `Comp.sub()` marks that point, the
  template function may fill the component in later;
* joins: `a <== b` / `a <-- b` whose right side is a single signal makes `a`
  an alias of `b` (one witness entry).  Any other right side (expression or
  constant) gives `a` its own entry.  Signals of the main component are never
  aliased away (they head the witness: outputs, then inputs).

Witness order (cf. snarkjs `.wtns`): w[0] = 1, then the
main component depth-first.
"""
from .field import P


class ConstraintError(Exception):
    """Mirrors circom_runtime's "Constraint doesn't match" error."""


class Sig:
    __slots__ = ("val", "alias", "idx")

    def __init__(self):
        self.val = None
        self.alias = None
        self.idx = -1

    def root(self):
        s = self
        while s.alias is not None:
            s = s.alias
        return s

    @property
    def v(self):
        s = self
        while s.alias is not None:
            s = s.alias
        if s.val is None:
            raise RuntimeError("signal read before assignment")
        return s.val


def val(x):
    """Value of a signal or of a plain integer expression."""
    if isinstance(x, Sig):
        return x.v
    return x % P


class Ctx:
    def __init__(self, sanity=True):
        self.sanity = sanity
        self.errors = []

    def fail(self, where, lhs, rhs):
        msg = "Constraint doesn't match %d != %d %s" % (lhs, rhs, where)
        self.errors.append(msg)
        if self.sanity:
            raise ConstraintError(msg)


class Comp:
    __slots__ = ("ctx", "name", "entries", "is_main")

    def __init__(self, ctx, name="main", is_main=False):
        self.ctx = ctx
        self.name = name
        self.entries = []
        self.is_main = is_main

    # -- structure ---------------------------------------------------------
    def sub(self, name):
        c = Comp(self.ctx, name)
        self.entries.append((name, c))
        return c

    def sig(self, name):
        s = Sig()
        self.entries.append((name, s))
        return s

    def sigs(self, name, n):
        return [self.sig("%s[%d]" % (name, i)) for i in range(n)]

    # -- assignment (dest <== x) --------------------------------------------
    def set(self, s, x):
        """`s <== x` (or `<--`) for an already declared signal."""
        if isinstance(x, Sig) and not self.is_main:
            s.alias = x
        else:
            s.val = val(x)
        return s

    def inp(self, name, x):
        """Declare a signal and assign it: input wired by the parent, or an
        own signal assigned once (`signal s; s <== x`)."""
        return self.set(self.sig(name), x)

    def inps(self, name, xs):
        return [self.inp("%s[%d]" % (name, i), x) for i, x in enumerate(xs)]

    # -- constraints ---------------------------------------------------------
    def check(self, lhs, rhs=0, what=""):
        lhs %= P
        rhs %= P
        if lhs != rhs:
            self.ctx.fail("%s %s" % (self.name, what), lhs, rhs)

    def check_bool(self, s, what="bool"):
        x = val(s)
        self.check(x * (x - 1), 0, what)


def number(main):
    """Assign witness indexes; returns the ordered list of root signals
    (index 0, the constant 1, is not in the list)."""
    order = []
    stack = [iter(main.entries)]
    while stack:
        try:
            _, obj = next(stack[-1])
        except StopIteration:
            stack.pop()
            continue
        if isinstance(obj, Comp):
            stack.append(iter(obj.entries))
        elif obj.alias is None:
            obj.idx = len(order) + 1
            order.append(obj)
    return order


def witness(main):
    """Full witness as a list of ints: [1, main depth-first ...]."""
    order = number(main)
    out = [1]
    for s in order:
        if s.val is None:
            raise RuntimeError("unassigned signal at witness index %d" % s.idx)
        out.append(s.val)
    return out


def symbols(main, prefix="main"):
    """name -> witness index (the `.sym` equivalent). Call after `number`."""
    table = {}

    def walk(c, pre):
        for name, obj in c.entries:
            full = pre + "." + name
            if isinstance(obj, Comp):
                walk(obj, full)
            else:
                table[full] = obj.root().idx
    walk(main, prefix)
    return table


def count(c):
    """Number of witness entries a component occupies (aliases excluded)."""
    n = 0
    stack = [c]
    while stack:
        x = stack.pop()
        for _, obj in x.entries:
            if isinstance(obj, Comp):
                stack.append(obj)
            elif obj.alias is None:
                n += 1
    return n
