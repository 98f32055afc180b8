function [features, validPoints] = extractFeatures(~, points, varargin)
% extractFeatures(I, points, "Method", "SIFT") (VO.m:83-84): the descriptors were computed together with
% the keypoints by detectSIFTFeatures.m; every point is valid.
features = points.Features_;
validPoints = SIFTPoints(points.Location, 'Scale', points.Scale, 'Orientation', points.Orientation, ...
                         'Metric', points.Metric, 'Octave', points.Octave, 'Layer', points.Layer);
end
