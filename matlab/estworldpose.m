function [worldPose, inlierIdx, status] = estworldpose(imagePoints, worldPoints, intrinsics, varargin)
% estworldpose on B200 (VO.m:123-127): P3P + MSAC.  intrinsics: cameraIntrinsics (FocalLength, PrincipalPoint).
K = [intrinsics.FocalLength, intrinsics.PrincipalPoint];
if nargout < 3
    [A, inlierIdx] = vo_p3p_mex(imagePoints, worldPoints, K, varargin{:});   % raises like the toolbox on failure
else
    [A, inlierIdx, status] = vo_p3p_mex(imagePoints, worldPoints, K, varargin{:});
end
worldPose = rigidtform3d(A);
end
