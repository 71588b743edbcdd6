"""TEST DOUBLE standing where the checkout's own module would be.  If a test ends up here, ivosw.hook did NOT
substitute the module (exactly the failure VERDICT r1 found in the PYTHONPATH recipe)."""
IS_CHECKOUT_ORIGINAL = True


def __getattr__(name):
    raise AssertionError("the checkout's own module was imported instead of the ivosw drop-in (asked for %r)" % name)
