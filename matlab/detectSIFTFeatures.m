function pts = detectSIFTFeatures(I, varargin)
% detectSIFTFeatures on B200 (VO.m:79-80).  Detection and description are one fused device call: the
% descriptors ride along in the returned struct and extractFeatures.m hands them out.
[desc, loc, scale, orient, metric, octave, layer] = vo_sift_mex(I, varargin{:});
pts = struct('Location', loc, 'Scale', scale, 'Orientation', orient, 'Metric', metric, ...
             'Octave', octave, 'Layer', layer, 'Count', size(loc, 1), 'Features_', desc);
end
