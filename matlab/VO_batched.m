% VO_batched.m -- the reference's main loop (VO.m:64-232) on the batched gateway vo_frames_mex.
%
% Everything before the loop (VO.m:1-62: datastores cam0/cam1, calibration p1/p2, intrinsics_l, pose, all_poses)
% stays as it is.  The loop body -- SIFT on both images, the five matchFeatures calls, find_remaining_points,
% triangulate and estworldpose -- runs on the B200 for B frames per call; what remains in MATLAB is reading the
% images, the sequential pose chain (VO.m:130) and the plotting.
B = 32;                                            % frames per call (plus one halo frame that re-seeds the tracker)
sz = size(readimage(cam0, 1));
prevL = []; prevR = [];
for i0 = 1:B:n_frames
    idx = i0:min(i0 + B - 1, n_frames);
    L = zeros([sz numel(idx)], 'uint8'); R = L;
    for k = 1:numel(idx)                           % VO.m:71-76 (undistortImage is the identity for KITTI: no coefficients)
        L(:, :, k) = readimage(cam0, idx(k));
        R(:, :, k) = readimage(cam1, idx(k));
    end
    if ~isempty(prevL), L = cat(3, prevL, L); R = cat(3, prevR, R); first = idx(1) - 2; else, first = 0; end
    [relA, status] = vo_frames_mex(L, R, p1, p2, 'FirstFrame', first);   % VO.m:79-127 for every frame of the stack
    for k = 2:size(L, 3)
        if status(k) ~= 0, error('estworldpose failed at frame %d (status %d)', first + k, status(k)); end
        rel_pose = rigidtform3d(relA(:, :, k));
        pose = rigidtform3d(pose.A * rel_pose.A);  % VO.m:130
        all_poses = [all_poses
                     pose];                        % VO.m:133
    end
    prevL = L(:, :, end); prevR = R(:, :, end);
end
