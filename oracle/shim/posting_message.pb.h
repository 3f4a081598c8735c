// Build shim (test infrastructure): intentionally empty (posting.h includes it but uses nothing from it).
