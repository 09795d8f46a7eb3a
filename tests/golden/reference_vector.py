"""The reference's own known-answer vector for the single histogram filter, kept verbatim as test data.

Source: /root/reference/SuperTest+/SuperAlgorithm+/STPTestHistogram.cpp
  :44-57  input texture, a 3x3 neighbourhood of 4x4 maps (12x12 samples)
  :59-63  STPNearestNeighbourInformation{{4,4},{3,3},{12,12}},  :71 radius 2
  :77-94  expected ordered (item, weight) bins of pixels 0, 8 and 15
  :103    weights compared within 5 * FLT_EPSILON relative
  :125-129 radii that must raise STPNumericDomainError
"""
import numpy as np

DIMENSION = (4, 4)
NEIGHBOUR = (3, 3)
RADIUS = 2
TEXTURE = np.array([
    2, 0, 2, 0, 1, 1, 0, 3, 3, 2, 1, 2,
    0, 1, 2, 1, 1, 1, 3, 2, 2, 0, 0, 1,
    0, 1, 0, 2, 1, 3, 1, 1, 1, 2, 1, 0,
    3, 3, 0, 1, 1, 2, 2, 2, 2, 0, 0, 3,
    3, 2, 3, 3, 0, 3, 2, 2, 1, 0, 0, 3,
    2, 0, 0, 1, 2, 0, 2, 2, 0, 2, 0, 3,
    1, 2, 3, 0, 3, 2, 1, 2, 2, 3, 0, 2,
    1, 1, 0, 3, 0, 2, 1, 0, 3, 2, 2, 1,
    2, 2, 0, 1, 0, 2, 0, 0, 0, 0, 1, 1,
    0, 1, 3, 3, 0, 3, 3, 3, 1, 0, 1, 1,
    3, 1, 3, 0, 2, 1, 1, 0, 2, 2, 2, 0,
    2, 0, 1, 1, 2, 3, 1, 2, 3, 2, 0, 1,
], dtype=np.uint16).reshape(12, 12)

# pixel -> ordered [(item, weight)]
EXPECTED = {
    0: [(0, 0.24), (3, 0.24), (2, 0.28), (1, 0.24)],
    8: [(0, 0.36), (3, 0.24), (1, 0.16), (2, 0.24)],
    15: [(0, 0.32), (3, 0.2), (1, 0.12), (2, 0.36)],
}
WEIGHT_RTOL = 5 * float(np.finfo(np.float32).eps)
BAD_RADII = (0, 128, 3)

# Full 16-pixel output of the reference for this input (counts out of 25), produced by running the reference's own
# filter (oracle/_ref) in the survey container -- SURVEY.md Appendix A.6.
FULL_COUNTS = [
    [(0, 6), (3, 6), (2, 7), (1, 6)], [(0, 3), (3, 4), (2, 11), (1, 7)], [(0, 3), (3, 3), (2, 12), (1, 7)], [(0, 4), (3, 3), (2, 13), (1, 5)],
    [(0, 7), (3, 6), (1, 5), (2, 7)], [(0, 5), (3, 4), (1, 5), (2, 11)], [(0, 5), (3, 3), (1, 4), (2, 13)], [(0, 5), (3, 3), (1, 3), (2, 14)],
    [(0, 9), (3, 6), (1, 4), (2, 6)], [(0, 8), (3, 4), (1, 4), (2, 9)], [(0, 9), (3, 3), (1, 3), (2, 10)], [(0, 8), (3, 3), (1, 3), (2, 11)],
    [(0, 9), (3, 7), (1, 4), (2, 5)], [(0, 8), (3, 6), (1, 4), (2, 7)], [(0, 9), (3, 5), (1, 3), (2, 8)], [(0, 8), (3, 5), (1, 3), (2, 9)],
]
