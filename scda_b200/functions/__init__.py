"""Host-facing target / proposal functions of the detector, same module names as the
reference's `functions/` package; the work runs on the device."""
