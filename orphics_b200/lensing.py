"""orphics.lensing hot-path mirror: the Hu-Okamoto quadratic estimator (lensing.qest)."""
