function [indexPairs, matchMetric] = matchFeatures(features1, features2, varargin)
% matchFeatures on B200 (VO.m:87, 283, 293, 311, 323): exhaustive SSD match, MatchThreshold / MaxRatio /
% Unique as name-value pairs; indexPairs is P x 2 uint32, 1-based, ascending in column 1.
[indexPairs, matchMetric] = vo_match_mex(features1, features2, varargin{:});
end
