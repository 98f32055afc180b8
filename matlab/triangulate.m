function [worldPoints, reprojErr, valid] = triangulate(matchedPoints1, matchedPoints2, camProjection1, camProjection2)
% triangulate on B200 (VO.m:114-115, CreateLandmarksFromFeatures.m:7): N x 2 (or 1 x 2) points, 3 x 4
% projection matrices.  Passing all N correspondences at once replaces the reference's per-point loop.
[worldPoints, reprojErr, valid] = vo_triangulate_mex(matchedPoints1, matchedPoints2, camProjection1, camProjection2);
end
