"""No-op stub of matplotlib.pyplot (the image has no matplotlib; triton.testing.Mark._run calls figure / subplot / ax.plot /
ax.fill_between / legend / set_* / savefig / show unconditionally).  Every attribute is an object that can be called, indexed
and have further attributes taken, and always returns itself."""


class _Any:
    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return self

    def __getitem__(self, key):
        return self

    def __iter__(self):
        return iter((self, self))


_ANY = _Any()


def __getattr__(name):
    return _ANY
